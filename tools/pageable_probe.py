"""Pageable-result e2e (what R's own allocation gives the shim): ipr_idw_f64 at cfg 3 into a NEVER-TOUCHED numpy
array (a fresh one per call, as R allocates a new vector per call) and into an already-touched one, for the
combinations of IPR_PREFAULT (pre-faulting threads ahead of the drain) and IPR_THP (transparent-huge-page hint)."""
import gc
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from interpolater_b200 import engine, synthetic

cx, cy, sx, sy, obs = synthetic.synthetic_case(1000, 1000, 2000, 3650)
obs_f = np.asfortranarray(obs)
engine.idw_grid(cx[:6400], cy[:6400], sx, sy, obs_f[:256])       # warm the library
for prefault, thp in (("0", "0"), ("1", "0"), ("1", "1"), ("0", "0"), ("1", "0"), ("1", "1")):
    os.environ["IPR_PREFAULT"] = prefault
    os.environ["IPR_THP"] = thp
    fresh = []
    for _ in range(2):
        pg = np.empty((cx.size, 3650), dtype=np.float64, order="F")     # never touched
        t = time.perf_counter()
        engine.idw_grid(cx, cy, sx, sy, obs_f, out=pg, devices=[0])
        fresh.append((time.perf_counter() - t) * 1e3)
        chk = float(np.nansum(pg[::9973, ::7]))
        if _ == 0:
            del pg
            gc.collect()
    t = time.perf_counter()
    engine.idw_grid(cx, cy, sx, sy, obs_f, out=pg, devices=[0])
    touched = (time.perf_counter() - t) * 1e3
    del pg
    gc.collect()
    print("IPR_PREFAULT=%s IPR_THP=%s: never-touched %.0f %.0f ms, touched %.0f ms; checksum %.6f" %
          (prefault, thp, fresh[0], fresh[1], touched, chk), flush=True)
