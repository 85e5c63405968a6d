"""Ordinary kriging at the cfg-3 geometry: 1M cells x 2,000 stations, a few dates (python tools/perf_kriging.py [n_dates])."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from interpolater_b200 import _native, engine, synthetic

D = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cx, cy, sx, sy, obs = synthetic.synthetic_case(1000, 1000, 2000, D, na_frac=0.05)
obs = obs + np.where(np.isnan(obs), 0.0, np.random.default_rng(0).uniform(0, 0.5, obs.shape))
lib = _native.load()
plan = engine.Plan(0, cx, cy, sx, sy, D)
plan.set_obs(obs)
out = torch.empty((D, cx.size), dtype=torch.float64, device="cuda")
for name, code in (("exponential", 1), ("spherical", 2), ("gaussian", 3)):
    model = np.full(D, code, dtype=np.int32)
    nug = np.full(D, 2.0)
    sill = np.full(D, 40.0)
    rng = np.full(D, 60e3 if code != 3 else 8e3)
    err = _native.errbuf()
    for it in range(2):
        plan.profile(True)
        plan.profile_read()
        plan.timer_start()
        rc = lib.ipr_plan_kriging(plan._h, model.ctypes.data, nug.ctypes.data, sill.ctypes.data, rng.ctypes.data, 0, D,
                                  out.data_ptr(), cx.size, err, _native.ERRLEN)
        _native.check(rc, err)
        ms = plan.timer_stop()
        ph = plan.profile_read()
    print("%-12s %d dates: %.1f ms total = %.2f ms/date  (solve %.1f ms, cell x station pass %.1f ms, IDW fallback pass %.1f ms); "
          "mean %.3f" % (name, D, ms, ms / D, ph["fixup"], ph["contract"] , ms - ph["fixup"] - ph["contract"], float(out.mean())), flush=True)
