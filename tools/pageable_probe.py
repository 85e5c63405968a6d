"""Pageable-result e2e (what R's own allocation gives the shim): ipr_idw_f64 at cfg 3 into a never-touched and then
an already-touched numpy array, with the streaming-store drain copy (default) and with memcpy (IPR_NT_COPY=0 is
read once per process, so the script re-executes itself)."""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) == 1:
    for nt in ("1", "0", "1", "0"):
        subprocess.run([sys.executable, __file__, nt], env=dict(os.environ, IPR_NT_COPY=nt))
    sys.exit(0)
import numpy as np
from interpolater_b200 import engine, synthetic

cx, cy, sx, sy, obs = synthetic.synthetic_case(1000, 1000, 2000, 3650)
obs_f = np.asfortranarray(obs)
engine.idw_grid(cx[:6400], cy[:6400], sx, sy, obs_f[:256])       # warm the library
pg = np.empty((cx.size, 3650), dtype=np.float64, order="F")
ms = []
for i in range(5):
    t = time.perf_counter()
    engine.idw_grid(cx, cy, sx, sy, obs_f, out=pg, devices=[0])
    ms.append((time.perf_counter() - t) * 1e3)
print("IPR_NT_COPY=%s: first touch %.0f ms, then %s ms; checksum %.6f" %
      (sys.argv[1], ms[0], " ".join("%.0f" % x for x in ms[1:]), float(np.nansum(pg[::9973, ::7]))), flush=True)
