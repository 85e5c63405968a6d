"""cfg 5 (20 % NA): FP64 masked GEMM vs error-free int8 mask contraction (IPR_MASK_INT8).
   python tools/perf_int8_mask.py [modes, default "01"] [iterations, default 3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolater_b200 import engine, synthetic
modes = sys.argv[1] if len(sys.argv) > 1 else "01"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cx, cy, sx, sy, obs = synthetic.synthetic_case(1000, 1000, 2000, 3650, na_frac=0.2)
C, D = cx.size, 3650
out = torch.empty((D, C), dtype=torch.float64, device="cuda")
ref = None
for mode in modes:
    os.environ["IPR_MASK_INT8"] = mode
    plan = engine.Plan(0, cx, cy, sx, sy, D); plan.set_obs(obs)
    best = 1e9
    for _ in range(iters):
        plan.reset_cache(); plan.timer_start(); plan.idw(out.data_ptr(), 2.0); best = min(best, plan.timer_stop())
    plan.timer_start(); plan.idw(out.data_ptr(), 2.0); warm = plan.timer_stop()
    print(f"int8={mode}: {best:.1f} ms ({warm:.1f} ms with cached weights)  {C*D/best*1e-6:.2f} G cell-days/s  launches {plan.launch_count}", flush=True)
    sample = out[:64, ::997].clone()
    if ref is None: ref = sample
    else: print("max rel diff vs fp64 path:", float(((sample - ref).abs() / ref.abs().clamp_min(1e-300)).max()))
    plan.close()
