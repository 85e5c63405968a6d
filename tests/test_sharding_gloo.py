"""N>1 host path on CPU: two gloo ranks each take their date block (no data-path collective), and
the gathered result equals the single-rank result.  The per-rank compute is the oracle standing in
for the CUDA engine (this box has no GPU); what is under test is the sharding / gathering logic
that bench.py and the C++ scheduler share."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, n_dates, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from interpolater_b200 import sharding, synthetic
    from oracle import reference_port as ref

    cx, cy, sx, sy, obs = synthetic.synthetic_case(6, 5, 7, n_dates, na_frac=0.1, all_na_dates=(1,))
    b = sharding.date_block_bounds(n_dates, world)
    lo, hi = b[rank], b[rank + 1]
    block = ref.idw_reference(cx, cy, sx, sy, obs[lo:hi], 2.0) if hi > lo else np.empty((cx.size, 0))
    # timing-style reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, block))
    if rank == 0:
        full = np.empty((cx.size, n_dates))
        for lo_, hi_, blk in gathered:
            full[:, lo_:hi_] = blk
        want = ref.idw_reference(cx, cy, sx, sy, obs, 2.0)
        same = np.array_equal(np.isnan(full), np.isnan(want)) and np.allclose(full[~np.isnan(want)], want[~np.isnan(want)], rtol=1e-13)
        q.put((same, float(t.item()), [g[:2] for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_dates", [300, 100])
def test_two_rank_date_sharding(n_dates):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_dates % 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_dates, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, tmax, blocks = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and tmax == 2.0
    assert blocks[0][0] == 0 and blocks[-1][1] == n_dates and blocks[0][1] == blocks[1][0]
