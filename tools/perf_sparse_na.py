"""cfg-3 shape with a small NA fraction: sparse denominator correction vs masked GEMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolater_b200 import engine, synthetic
na = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
cx, cy, sx, sy, obs = synthetic.synthetic_case(1000, 1000, 2000, 3650, na_frac=na)
C, D, S = cx.size, 3650, 2000
out = torch.empty((D, C), dtype=torch.float64, device="cuda")
for mode in ("1", "0"):
    os.environ["IPR_SPARSE_NA"] = mode
    plan = engine.Plan(0, cx, cy, sx, sy, D); plan.set_obs(obs)
    best = 1e9
    for _ in range(3):
        plan.reset_cache(); plan.timer_start(); plan.idw(out.data_ptr(), 2.0); best = min(best, plan.timer_stop())
    print(f"NA {na:.3f} sparse={mode}: {best:.1f} ms  {C*D/best*1e-6:.2f} G cell-days/s  launches {plan.launch_count}", flush=True)
    plan.close()
