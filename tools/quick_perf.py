"""Mid-size throughput probe through the plan API (device-resident), for kernel tuning."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from interpolater_b200 import engine, synthetic

ncol, nrow, S, D = 296, 256, 2000, 3650
which = sys.argv[1] if len(sys.argv) > 1 else "idw,cressman,masked"
cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, S, D)
C = cx.size
plan = engine.Plan(0, cx, cy, sx, sy, D)
plan.set_obs(obs)
tag = os.environ.get("IPR_WCACHE_MB", "all")
def run(name, fn, flops, reps=3):
    best = 1e9
    for _ in range(reps):
        if os.environ.get("IPR_REGEN"): plan.reset_cache()
        plan.timer_start(); fn(); best = min(best, plan.timer_stop())
    print(f"[wcache={tag}] {name}: {best:.2f} ms {C*D/best*1e-6:.3f} Gcd/s {flops/best*1e-9:.2f} TF ({flops/best*1e-9/37.12*100:.1f}%)", flush=True)
if "idw" in which:
    out = torch.empty((D, C), dtype=torch.float64, device="cuda")
    run("IDW", lambda: plan.idw(out.data_ptr(), 2.0), 2.0*C*S*D)
if "cressman" in which:
    outc = torch.empty((4, D, C), dtype=torch.float64, device="cuda")
    run("Cressman4", lambda: plan.cressman(outc.data_ptr(), np.array([25e3,50e3,100e3,200e3])), 8.0*C*S*D, 2)
if "masked" in which:
    obs3 = obs.copy(); obs3[np.random.default_rng(5).random(obs3.shape) < 0.2] = np.nan
    plan.set_obs(obs3)
    out = torch.empty((D, C), dtype=torch.float64, device="cuda")
    run("IDW-masked", lambda: plan.idw(out.data_ptr(), 2.0), 4.0*C*S*D, 2)
ex, de = plan.stage_blocks()
print(f"stage blocks executed/dense = {ex}/{de} = {ex/max(de,1):.3f}")
