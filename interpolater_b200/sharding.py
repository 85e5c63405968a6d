"""Sharding: every output element depends only on its own cell and date, so GPUs take
disjoint pieces with no exchange step — independent date blocks when every GPU owns its own stretch of
the record (`date_block_bounds`, the weak-scaling layout of bench.py), contiguous cell ranges when one
job is split (`cell_block_bounds`)."""
from __future__ import annotations


def cell_block_bounds(n_cells: int, n_shards: int, align: int = 64):
    """The split the C++ scheduler (csrc/ipr_api.cu, run_host_job) and `bench.py --scaling strong` use
    when ONE job is spread over several GPUs: contiguous cell ranges on 64-cell tile boundaries.
    Nothing is computed twice (a date split would make every GPU regenerate all cell x station
    weights) and the shares are equal to within one tile.  Returns n_shards+1 monotone bounds."""
    if n_cells < 0 or n_shards < 1 or align < 1:
        raise ValueError("n_cells >= 0, n_shards >= 1, align >= 1 required")
    n_blk = (n_cells + align - 1) // align
    b = [min(n_cells, (n_blk * g // n_shards) * align) for g in range(n_shards + 1)]
    b[n_shards] = n_cells
    return b


def date_block_bounds(n_dates: int, n_shards: int, align: int = 32):
    """Returns n_shards+1 monotone bounds; shard g owns dates [bounds[g], bounds[g+1])."""
    if n_dates < 0 or n_shards < 1 or align < 1:
        raise ValueError("n_dates >= 0, n_shards >= 1, align >= 1 required")
    n_blk = (n_dates + align - 1) // align
    b = [min(n_dates, (n_blk * g // n_shards) * align) for g in range(n_shards + 1)]
    b[n_shards] = n_dates
    return b
