"""Phase timing of the host-buffer entry point (where does e2e time go?)."""
import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from interpolater_b200 import engine, synthetic, _native
lib = _native.load()
ncol, nrow, S, D = 1000, 1000, 2000, 3650
cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, S, D)
Cn = cx.size
t = time.perf_counter(); plan = engine.Plan(0, cx, cy, sx, sy, D); torch.cuda.synchronize(); print("plan create", time.perf_counter() - t)
t = time.perf_counter(); plan.set_obs(obs); plan.sync(); print("set_obs", time.perf_counter() - t)
t = time.perf_counter(); plan.close(); print("plan destroy", time.perf_counter() - t)
t = time.perf_counter(); plan = engine.Plan(0, cx, cy, sx, sy, D); torch.cuda.synchronize(); print("plan create (2nd)", time.perf_counter() - t)
plan.close()
hp = lib.ipr_host_alloc(Cn * D * 8)
out = np.ctypeslib.as_array((C.c_double * (Cn * D)).from_address(hp)).reshape(D, Cn).T
obs_f = np.asfortranarray(obs)
for i in range(12):
    t = time.perf_counter(); engine.idw_grid(cx, cy, sx, sy, obs_f, out=out, devices=[0]); print("idw_grid e2e", time.perf_counter() - t, flush=True)
