"""One Cressman call per radius given on the command line (metres), cfg-4 geometry, for ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from interpolater_b200 import engine, synthetic

radii = [float(x) for x in sys.argv[1:]] or [200e3]
D = int(os.environ.get("DATES", "3650"))
cx, cy, sx, sy, obs = synthetic.synthetic_case(1000, 1000, 2000, D)
plan = engine.Plan(0, cx, cy, sx, sy, D)
plan.set_obs(obs)
out = torch.empty((1, D, cx.size), dtype=torch.float64, device="cuda")
for r in radii:
    plan.cressman(out.data_ptr(), np.array([r]))
plan.sync()
print("ok", float(torch.nan_to_num(out[0, 0]).sum()))
