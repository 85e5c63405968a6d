"""cfg 4 (Cressman, 4 radii) device-resident step with per-radius phase timings.
IPR_CRESSMAN_LEGACY=1 python tools/perf_cressman.py   -> the round-1 path for comparison."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from interpolater_b200 import engine, synthetic

ncol = nrow = 1000
S, D = 2000, int(sys.argv[1]) if len(sys.argv) > 1 else 3650
radii = np.array([float(x) for x in os.environ.get("RADII", "25e3,50e3,100e3,200e3").split(",")])
cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, S, D)
plan = engine.Plan(0, cx, cy, sx, sy, D)
d_obs = torch.from_numpy(np.asfortranarray(obs).T.copy()).cuda()
out = torch.empty((len(radii), D, cx.size), dtype=torch.float64, device="cuda")
for it in range(3):
    plan.reset_cache()
    plan.set_obs_device(d_obs.data_ptr())
    ex0, de0 = plan.stage_blocks()
    plan.profile(True)
    plan.profile_read()
    plan.timer_start()
    plan.cressman(out.data_ptr(), radii)
    ms = plan.timer_stop()
    ph = plan.profile_read()
    ex1, de1 = plan.stage_blocks()
    print("legacy" if os.environ.get("IPR_CRESSMAN_LEGACY") == "1" else "fused", "step %.2f ms  weights %.2f  radii %s  executed %.4f of dense"
          % (ms, ph["weights"], ["%.2f" % x for x in ph["radius"][:len(radii)]], (ex1 - ex0) / max(de1 - de0, 1)), flush=True)
print("NA fraction per radius:", [float(torch.isnan(out[i, 0]).double().mean()) for i in range(len(radii))])
