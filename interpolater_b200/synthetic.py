"""Deterministic synthetic workloads (SURVEY §8d cfg 3-5): regular grid, uniformly scattered
stations, zero-inflated 'precipitation'.  Used by bench.py and by the tests so that both sides of a
parity check see the same inputs.  Pure NumPy host code; not on the hot path."""
from __future__ import annotations

import numpy as np


def regular_grid(ncol, nrow, res_m=1000.0, xmin=200000.0, ymin=9000000.0):
    """Cell centres in terra order (row-major from the top-left), R/IDW.R:249-257."""
    xc = xmin + (np.arange(ncol, dtype=np.float64) + 0.5) * res_m
    yc = (ymin + nrow * res_m) - (np.arange(nrow, dtype=np.float64) + 0.5) * res_m
    return np.tile(xc, nrow), np.repeat(yc, ncol)


def synthetic_case(ncol, nrow, n_st, n_dates, res_m=1000.0, xmin=200000.0, ymin=9000000.0, na_frac=0.0,
                   plant_zero_dist=0, all_na_dates=(), all_zero_dates=(), seed=20260101):
    """Returns (cell_x, cell_y, st_x, st_y, obs[n_dates, n_st]).  Wet w.p. 0.45, amount
    Gamma(0.8, 8) rounded to 0.1; `plant_zero_dist` stations sit exactly on cell centres."""
    cx, cy = regular_grid(ncol, nrow, res_m, xmin, ymin)
    rng = np.random.default_rng(seed)
    sx = xmin + rng.random(n_st) * ncol * res_m
    sy = ymin + rng.random(n_st) * nrow * res_m
    if plant_zero_dist:
        idx = rng.choice(ncol * nrow, size=plant_zero_dist, replace=False)
        sx[:plant_zero_dist] = cx[idx]
        sy[:plant_zero_dist] = cy[idx]
    rng2 = np.random.default_rng(seed + 1)
    wet = rng2.random((n_dates, n_st)) < 0.45
    amt = np.round(rng2.gamma(0.8, 8.0, size=(n_dates, n_st)), 1)
    obs = np.where(wet, amt, 0.0)
    if na_frac > 0:
        rng3 = np.random.default_rng(seed + 2)
        obs[rng3.random((n_dates, n_st)) < na_frac] = np.nan
    for t in all_na_dates:
        obs[t, :] = np.nan
    for t in all_zero_dates:
        obs[t, :] = np.where(np.isnan(obs[t, :]), np.nan, 0.0)
    return cx, cy, sx, sy, obs
