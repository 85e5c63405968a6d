"""Where does the host-buffer entry point spend its time?  (python tools/e2e_trace.py [n_calls])

Runs cfg 3 (1M cells x 2,000 stations x 3,650 days) through ipr_idw_f64 with IPR_TRACE=1 so that every call
prints its phase timestamps to stderr; stdout gets one line per call.  Also times what an R caller would
pay for the RESULT buffer itself: cudaHostAlloc of 29.2 GB, a pageable buffer fresh / touched, with and
without the transparent-huge-page hint."""
import ctypes as C
import os
import sys
import time

os.environ.setdefault("IPR_TRACE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from interpolater_b200 import _native, engine, synthetic  # noqa: E402

lib = _native.load()
n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ncol, nrow, S, D = 1000, 1000, 2000, 3650
cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, S, D)
Cn = cx.size
obs_f = np.asfortranarray(obs)
try:
    print("THP:", open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip(), flush=True)
except OSError:
    pass
t = time.perf_counter()
hp = lib.ipr_host_alloc(Cn * D * 8)
print("cudaHostAlloc 29.2 GB: %.3f s" % (time.perf_counter() - t), flush=True)
out = np.ctypeslib.as_array((C.c_double * (Cn * D)).from_address(hp)).reshape(D, Cn).T
for pace in ("1", "0", "1", "0"):
    os.environ["IPR_PACE"] = pace       # 1: host-paced ring (progress / cancel per chunk); 0: everything queued at once
    for i in range(n_calls):
        t = time.perf_counter()
        engine.idw_grid(cx, cy, sx, sy, obs_f, out=out, devices=[0])
        print("pinned IPR_PACE=%s call %d: %.1f ms" % (pace, i, (time.perf_counter() - t) * 1e3), flush=True)
        print("---- pace", pace, file=sys.stderr, flush=True)
os.environ.pop("IPR_PACE")
t = time.perf_counter()
lib.ipr_host_free(hp)
print("cudaFreeHost 29.2 GB: %.3f s" % (time.perf_counter() - t), flush=True)
os.environ.pop("IPR_TRACE", None)
for thp in ("0",):
    os.environ["IPR_THP"] = thp
    pg = np.empty((Cn, D), dtype=np.float64, order="F")      # fresh, never touched: what allocVector returns
    for i in range(3):
        t = time.perf_counter()
        engine.idw_grid(cx, cy, sx, sy, obs_f, out=pg, devices=[0])
        print("pageable IPR_THP=%s call %d (%s): %.1f ms" % (thp, i, "fresh" if i == 0 else "touched",
                                                              (time.perf_counter() - t) * 1e3), flush=True)
    del pg
