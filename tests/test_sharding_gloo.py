"""N>1 host path, world size 2 on the gloo backend (runs on CPU).

What is under test is what bench.py's strong-scaling arm and the C++ scheduler share: the cell ranges
come from the LIBRARY (`ipr_cell_split_bounds`, the very function `run_host_job` splits a job with), every
rank fills only its own columns of every layer, timings are reduced with MAX over ranks, and no rank needs
anything from another one (no data-path collective; the gather below exists only so that rank 0 can check
the result).  On CPU the per-cell payload is a closed-form function of (cell, date), so that a wrong,
overlapping or missing range shows up as a wrong value; the `gpu`-marked variant runs the real engine on
each rank's range and compares the assembled matrix with the one-call result bit by bit."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _payload(cells, n_dates):
    return cells[:, None] * 4096.0 + np.arange(n_dates, dtype=np.float64)[None, :]


def _worker(rank, world, port, shape, use_engine, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from interpolater_b200 import engine, sharding, synthetic

    ncol, nrow, n_st, n_dates = shape
    cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, n_st, n_dates, na_frac=0.1, all_na_dates=(1,))
    b = engine.cell_split_bounds(cx.size, world)              # the C++ scheduler's own split
    assert b == sharding.cell_block_bounds(cx.size, world)    # bench.py's strong arm uses this twin
    lo, hi = b[rank], b[rank + 1]
    if use_engine:
        dev = rank % max(1, torch.cuda.device_count())
        block = engine.idw_grid(cx[lo:hi], cy[lo:hi], sx, sy, obs, devices=[dev]) if hi > lo else np.empty((0, n_dates))
    else:
        block = _payload(np.arange(lo, hi, dtype=np.float64), n_dates)
    # timing-style reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, block))
    if rank == 0:
        full = np.full((cx.size, n_dates), -1.0)
        covered = np.zeros(cx.size, dtype=np.int64)
        for lo_, hi_, blk in gathered:
            full[lo_:hi_, :] = blk
            covered[lo_:hi_] += 1
        if use_engine:
            want = engine.idw_grid(cx, cy, sx, sy, obs, devices=[0])
            same = bool(np.array_equal(np.isnan(full), np.isnan(want)) and
                        np.allclose(full[~np.isnan(want)], want[~np.isnan(want)], rtol=1e-12, atol=0))
        else:
            same = bool(np.array_equal(full, _payload(np.arange(cx.size, dtype=np.float64), n_dates)))
        q.put((same, bool(np.all(covered == 1)), float(t.item()), [g[:2] for g in gathered]))
    dist.barrier()
    dist.destroy_process_group()


def _run(shape, use_engine):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + shape[0] % 7 + (11 if use_engine else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, shape, use_engine, q)) for r in range(2)]
    for p in procs:
        p.start()
    same, covered_once, tmax, blocks = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert same and covered_once and tmax == 2.0
    n_cells = shape[0] * shape[1]
    assert blocks[0][0] == 0 and blocks[-1][1] == n_cells and blocks[0][1] == blocks[1][0]
    assert blocks[0][1] % 64 == 0 or blocks[0][1] == n_cells      # 64-cell tile boundary


@pytest.mark.parametrize("shape", [(40, 30, 7, 50), (9, 7, 5, 20), (8, 8, 3, 4)])
def test_two_rank_cell_sharding(shape):
    _run(shape, use_engine=False)


@pytest.mark.gpu
def test_two_rank_cell_sharding_with_the_engine():
    _run((60, 45, 40, 300), use_engine=True)
