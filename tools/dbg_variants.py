import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from interpolater_b200 import engine, synthetic, _native
ncol, nrow, S, D = 296, 256, 2000, 3584
cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, S, D)
C_ = cx.size
plan = engine.Plan(0, cx, cy, sx, sy, D); plan.set_obs(obs)
lib = _native.load()
lib.ipr_debug_variant.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_char_p, C.c_size_t]
out = torch.empty((D, C_), dtype=torch.float64, device="cuda")
names = {0: "full", 1: "no-LDS", 2: "no-refill", 3: "no-LDS+no-refill", 4: "no-weightgen", 7: "DMMA only"}
for dbg in (0, 1, 2, 3, 4, 7):
    best = 1e9
    for _ in range(3):
        err = _native.errbuf()
        plan.timer_start(); rc = lib.ipr_debug_variant(plan._h, dbg, C.c_void_p(out.data_ptr()), err, 512); ms = plan.timer_stop()
        assert rc == 0, err.value
        best = min(best, ms)
    fl = 2.0 * C_ * S * D
    print(f"dbg={dbg} {names[dbg]:18s}: {best:.2f} ms {fl/best*1e-9:.2f} TF ({fl/best*1e-9/37.12*100:.1f}%)", flush=True)
