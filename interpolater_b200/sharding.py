"""Date-block sharding: every output element depends only on its own cell and date, so
GPUs take contiguous date blocks with no exchange step.  Same rule as the C++ scheduler in
csrc/ipr_api.cu (run_host_job): blocks as even as possible in units of `align` = 32 dates (the
width step of the contraction's tail tiles; each device numbers its own dates from 0, so no
coarser alignment is needed)."""
from __future__ import annotations


def date_block_bounds(n_dates: int, n_shards: int, align: int = 32):
    """Returns n_shards+1 monotone bounds; shard g owns dates [bounds[g], bounds[g+1])."""
    if n_dates < 0 or n_shards < 1 or align < 1:
        raise ValueError("n_dates >= 0, n_shards >= 1, align >= 1 required")
    n_blk = (n_dates + align - 1) // align
    b = [min(n_dates, (n_blk * g // n_shards) * align) for g in range(n_shards + 1)]
    b[n_shards] = n_dates
    return b
