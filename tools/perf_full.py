"""Full-size (cfg 3/4) device-resident timing through the plan API, for A/B comparisons."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from interpolater_b200 import engine, synthetic
ncol, nrow, S, D = 1000, 1000, 2000, 3650
cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, S, D)
C = cx.size
plan = engine.Plan(0, cx, cy, sx, sy, D); plan.set_obs(obs)
tag = os.path.basename(os.environ.get("IPR_LIB_OVERRIDE", "default"))
out = torch.empty((D, C), dtype=torch.float64, device="cuda")
def run(name, fn, reps=3):
    best = 1e9
    for _ in range(reps):
        plan.reset_cache(); plan.timer_start(); fn(); best = min(best, plan.timer_stop())
    return best
ms = run("idw", lambda: plan.idw(out.data_ptr(), 2.0))
print(f"[{tag}] IDW {ms:.1f} ms  {2.0*C*S*D/ms*1e-9/37.12*100:.1f}% of peak", flush=True)
for R in (25e3, 50e3, 100e3, 200e3):
    e0, d0 = plan.stage_blocks()
    ms = run("c", lambda: plan.cressman(out.data_ptr(), np.array([R])), reps=2)
    e1, d1 = plan.stage_blocks()
    print(f"[{tag}] Cressman R={R/1e3:.0f} km: {ms:.1f} ms  executed blocks {(e1-e0)/(d1-d0):.4f}", flush=True)
