"""One process driving n GPUs (the device list of ipr_idw_f64): where does the time go?
(python tools/inproc_probe.py [n_gpus])

1. raw copy probe: every GPU copies a (128 x its cell range) block per step into ITS COLUMN RANGE of one pinned
   host matrix (the strided 2-D copy the engine issues), against the same bytes into a private contiguous
   pinned buffer — is the pitched D2H itself slower than the contiguous one?
2. the engine call with IPR_TRACE=1 (phase timestamps of every device thread on stderr)."""
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from interpolater_b200 import _native, engine, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
C, D = 1_000_000, 3650
per = C // n
host = torch.empty((D, C), dtype=torch.float64, pin_memory=True)
priv = [torch.empty((128, per), dtype=torch.float64, pin_memory=True) for _ in range(n)]
dev = [torch.zeros((128, per), dtype=torch.float64, device=f"cuda:{g}") for g in range(n)]
streams = [torch.cuda.Stream(device=g) for g in range(n)]


def run(kind):
    def work(g):
        with torch.cuda.device(g), torch.cuda.stream(streams[g]):
            for t0 in range(0, D - 127, 128):
                if kind == "pitched":
                    host[t0:t0 + 128, g * per:(g + 1) * per].copy_(dev[g], non_blocking=True)
                else:
                    priv[g].copy_(dev[g], non_blocking=True)
            streams[g].synchronize()
    th = [threading.Thread(target=work, args=(g,)) for g in range(n)]
    t = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    s = time.perf_counter() - t
    nbytes = n * per * 8 * 128 * ((D - 127 + 127) // 128)
    return nbytes / s * 1e-9


for kind in ("contiguous", "pitched", "contiguous", "pitched"):
    print(f"{n} GPU(s), {kind} D2H from one process: {run(kind):.1f} GB/s aggregate", flush=True)
del priv, dev

os.environ["IPR_TRACE"] = "1"
cx, cy, sx, sy, obs = synthetic.synthetic_case(1000, 1000, 2000, D)
obs_f = np.asfortranarray(obs)
out = host.numpy().T          # (C, D) Fortran-ordered view of the pinned matrix
# IPR_D2H_ROWS is read once per process: the row-wise variant is run by re-executing this script with the switch set
mode = os.environ.get("IPR_D2H_ROWS", "1 (default)")
for i in range(4):
    t = time.perf_counter()
    engine.idw_grid(cx, cy, sx, sy, obs_f, out=out, devices=list(range(n)))
    print("ipr_idw_f64(devices=0..%d, IPR_D2H_ROWS=%s) call %d: %.1f ms" % (n - 1, mode, i, (time.perf_counter() - t) * 1e3), flush=True)
    print("---- call", i, file=sys.stderr, flush=True)
