"""ORACLE / TEST INFRASTRUCTURE — not product code.  PARITY UNPINNED (see below).

CPU restatement (NumPy float64) of the reference's IDW / Cressman hot path and of the
validation code that consumes its output.  Every function cites the reference lines it follows
(paths are under /root/reference).  The restatement is deliberately *literal*: it materialises
the dense cell x station distance and weight matrices and loops over dates exactly like the R
code, including the quirks a "clean" implementation would not have:

  * IDW early-outs (all-NA date -> NA layer; all |obs| < 1e-12 -> exact-zero layer),
    R/IDW.R:294-302;
  * IDW zero-distance running mean over co-located available stations, R/IDW.R:310-326;
  * Cressman's denominator sums the weights of ALL stations, including those that are NA on
    the date (R/Cressman.R:315), so an all-NA date gives 0, not NA, where any station is in range;
  * `!is.finite(x) | denom == 0 -> NA`, R/IDW.R:332,339 and R/Cressman.R:317.

Parity status: **unpinned**.  The reference's own tests hold no numeric golden vectors
(tests/testthat/test-IDW.R, test-Cressman.R assert class / nlyr / error strings only) and R is not
installed in the build container or on the GPU box (probed: no Rscript), so this oracle could not
be checked against the real package.  Arithmetic that lives in un-vendored dependencies is
restated from their documented behaviour: terra (unpinned, DESCRIPTION:21) `rast(ext, resolution)`,
`as.data.frame(xy=TRUE)`, planar `distance`, `extract`; R >= 4.4 `^` (libm pow), `%*%` (dgemv),
`rowSums` (long double accumulation), `round` (R >= 4.0.0 algorithm).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this module, and only as the checker / the timed CPU baseline.
"""
from __future__ import annotations

import math
import struct

import numpy as np

#: R's NA_real_ bit pattern (a quiet NaN with low word 1954)
NA_REAL_BITS = 0x7FF00000000007A2
NA_REAL = struct.unpack("<d", struct.pack("<Q", NA_REAL_BITS))[0]


# --------------------------------------------------------------------------------------
# grid (terra::rast(ext, resolution) + as.data.frame(xy=TRUE, cells=TRUE)), R/IDW.R:198-209,
# :249-257; R/Cressman.R:216-227, :268-277
# --------------------------------------------------------------------------------------
def grid_from_extent(xmin, xmax, ymin, ymax, res_m):
    """terra `rast(ext, resolution=)`: the extent keeps (xmin, ymin); ncol/nrow are
    round(range/res) (>= 1) and xmax/ymax are moved to xmin + ncol*res, ymin + nrow*res.
    Cells are numbered row-major from the TOP-left; centre x = xmin + (col+.5)*res,
    y = ymax' - (row+.5)*res.  Returns dict(ncol, nrow, xmin, xmax, ymin, ymax, X, Y)."""
    ncol = int(max(1.0, math.floor((xmax - xmin) / res_m + 0.5)))
    nrow = int(max(1.0, math.floor((ymax - ymin) / res_m + 0.5)))
    xmax2 = xmin + ncol * res_m
    ymax2 = ymin + nrow * res_m
    xr = (xmax2 - xmin) / ncol
    yr = (ymax2 - ymin) / nrow
    cols = np.arange(ncol, dtype=np.float64)
    rows = np.arange(nrow, dtype=np.float64)
    xc = xmin + (cols + 0.5) * xr
    yc = ymax2 - (rows + 0.5) * yr
    X = np.tile(xc, nrow)
    Y = np.repeat(yc, ncol)
    return dict(ncol=ncol, nrow=nrow, xmin=xmin, xmax=xmax2, ymin=ymin, ymax=ymax2, X=X, Y=Y)


def distance_matrix(cx, cy, sx, sy):
    """terra::distance(points, points) for a projected CRS: planar sqrt(dx^2 + dy^2),
    n_cells x n_est (R/IDW.R:260-268, R/Cressman.R:279-288)."""
    dx = np.asfortranarray(cx[:, None] - sx[None, :])   # column-major like an R matrix
    dy = np.asfortranarray(cy[:, None] - sy[None, :])
    return np.sqrt(dx * dx + dy * dy)


def _is_na(v):
    """R's is.na() on doubles: TRUE for NA_real_ and for NaN."""
    return np.isnan(v)


# --------------------------------------------------------------------------------------
# IDW  (R/IDW.R:249-355)
# --------------------------------------------------------------------------------------
def idw_setup(cx, cy, sx, sy, p=2.0):
    """Date-independent part of IDW(): dist_mat (R/IDW.R:260-268) and w_base = dist^-p (:271-278)."""
    cx = np.asarray(cx, np.float64)
    cy = np.asarray(cy, np.float64)
    sx = np.asarray(sx, np.float64)
    sy = np.asarray(sy, np.float64)
    dist_mat = distance_matrix(cx, cy, sx, sy)                      # :260-268
    with np.errstate(divide="ignore", over="ignore"):
        w_base = np.power(dist_mat, -p)                             # :273  (d == 0 -> Inf)
    return dist_mat, w_base


def idw_one_date(dist_mat, w_base, obs_vec):
    """idw_predict(data_obs) for one date (R/IDW.R:288-342), statement by statement.
    Returns the n_cells vector of the date's layer (NaN where the reference yields NA)."""
    n_cells = dist_mat.shape[0]
    not_na = ~_is_na(obs_vec)                                       # :290
    if not_na.any():                                                # :294
        if np.all(np.abs(obs_vec[not_na]) < 1e-12):                 # :295-296
            return np.zeros(n_cells)                                # :297  r_zero
    else:
        return np.full(n_cells, np.nan)                             # :301  r_na
    has_obs = not_na                                                # :305
    w = w_base[:, has_obs]                                          # :306
    o = obs_vec[has_obs]                                            # :307
    d_sub = dist_mat[:, has_obs]
    zero_cols = np.nonzero((d_sub == 0).sum(axis=0) > 0)[0]         # :310
    if zero_cols.size > 0:                                          # :312
        pred = np.full(n_cells, np.nan)                             # :313
        hit_count = np.zeros(n_cells, np.int64)                     # :314
        for k in zero_cols:                                         # :315
            idx = np.nonzero(d_sub[:, k] == 0)[0]                   # :316-317
            if idx.size:                                            # :318
                cur = pred[idx]
                with np.errstate(invalid="ignore"):
                    upd = (cur * hit_count[idx] + o[k]) / (hit_count[idx] + 1)   # :322
                pred[idx] = np.where(np.isnan(cur), o[k], upd)      # :319-323
                hit_count[idx] += 1                                 # :324
        need_idw = np.nonzero(np.isnan(pred))[0]                    # :327
        if need_idw.size:                                           # :328
            wn = w[need_idw]
            with np.errstate(invalid="ignore", over="ignore"):
                numer = wn @ o                                      # :329  dgemv
                denom = _row_sums_na_rm(wn)                         # :330
            with np.errstate(invalid="ignore", divide="ignore"):
                vals = numer / denom                                # :331
            vals[~np.isfinite(vals) | (denom == 0)] = np.nan        # :332
            pred[need_idw] = vals                                   # :333
    else:
        with np.errstate(invalid="ignore", over="ignore"):
            numer = w @ o                                           # :336
            denom = _row_sums_na_rm(w)                              # :337
        with np.errstate(invalid="ignore", divide="ignore"):
            pred = numer / denom                                    # :338
        pred[~np.isfinite(pred) | (denom == 0)] = np.nan            # :339
    return pred


def idw_reference(cx, cy, sx, sy, obs, p=2.0):
    """obs: (n_dates, n_est) with NaN for missing.  Returns (n_cells, n_dates) float64 with NaN
    where the reference yields NA: the date loop of R/IDW.R:352 over idw_one_date."""
    obs = np.asarray(obs, np.float64)
    dist_mat, w_base = idw_setup(cx, cy, sx, sy, p)
    out = np.empty((dist_mat.shape[0], obs.shape[0]), np.float64)
    for t in range(obs.shape[0]):                                   # :352  pblapply over dates
        out[:, t] = idw_one_date(dist_mat, w_base, obs[t])
    return out


#: rowSums accumulates in long double in R; emulating that in NumPy costs an extra widened copy.
#: bench.py's CPU-baseline leg sets this to False (plain float64 pairwise sums, ~1e-16 apart) so
#: that the timed port is not slower than R's C loop would be.
# --------------------------------------------------------------------------------------
# Kriging_Ordinary: the IDW fallback of perform_kriging and its per-date prelude
# --------------------------------------------------------------------------------------
def kriging_idw_fallback_one_date(dist_pred_obs, obs_vec):
    """One date of perform_kriging when `use_idw` is TRUE (flat variogram, or a singular / ill-conditioned
    gamma matrix), R/Kriging_Ordinary.R:405-423 and :465-476:

        available_values <- obs_values[!is.na(obs_values)]
        if (length(available_stations) < 2)  value = mean(available_values), or 0 when there is none
        if (length(unique(available_values)) == 1)  value = that value
        distances[distances == 0] <- 1e-10 ; weights <- 1 / (distances^2)
        sum(weights * available_values) / sum(weights)

    dist_pred_obs is the cell x station distance matrix (terra::distance, :399-402)."""
    obs_vec = np.asarray(obs_vec, dtype=np.float64)
    avail = ~_is_na(obs_vec)
    vals = obs_vec[avail]
    n_cells = dist_pred_obs.shape[0]
    if vals.size < 2:                                             # :410-417
        return np.full(n_cells, float(vals.mean()) if vals.size > 0 else 0.0)
    if np.unique(vals).size == 1:                                 # :420-423
        return np.full(n_cells, float(vals[0]))
    d = dist_pred_obs[:, avail].copy()
    d[d == 0] = 1e-10                                             # :470
    w = 1.0 / (d * d)                                             # :471 (R evaluates x^2 as x * x)
    return (w * vals[None, :]).sum(axis=1) / w.sum(axis=1)       # :472


VARIOGRAM_MODELS = {"exponential": 1, "spherical": 2, "gaussian": 3, "linear": 4}


def variogram_models(h, nugget, sill, rng, model):
    """variogram_models(), R/Kriging_Ordinary.R:262-275 (model by name or by the engine's code 1..4)."""
    if not isinstance(model, str):
        model = {v: k for k, v in VARIOGRAM_MODELS.items()}[int(model)]
    h = np.asarray(h, dtype=np.float64)
    if model == "exponential":
        return nugget + sill * (1 - np.exp(-3 * h / rng))
    if model == "spherical":
        return np.where(h <= rng, nugget + sill * (1.5 * h / rng - 0.5 * (h / rng) ** 3), nugget + sill)
    if model == "gaussian":
        return nugget + sill * (1 - np.exp(-3 * (h / rng) ** 2))
    if model == "linear":
        return np.where(h <= rng, nugget + sill * h / rng, nugget + sill)
    raise ValueError(model)


def kriging_one_date(dist_pred_obs, dist_obs_obs, obs_vec, params):
    """perform_kriging for one date, R/Kriging_Ordinary.R:405-508.  params: None or a dict with use_idw=True ->
    the IDW fallback; else dict(nugget, sill, range, model).  The kriging branch is the literal per-cell
    `solve(gamma_matrix, gamma_vector)` (:478-497): gamma_matrix is initialised with the nugget (so the diagonal
    holds it), off-diagonal entries are the variogram of the station distances, the last row / column are 1 and
    the corner 0 (:431-452); prediction max(0, sum(weights[1:n] * values)), mean(values) where solve() fails."""
    obs_vec = np.asarray(obs_vec, dtype=np.float64)
    if params is None or params.get("use_idw"):
        return kriging_idw_fallback_one_date(dist_pred_obs, obs_vec)
    avail = ~_is_na(obs_vec)
    vals = obs_vec[avail]
    n_cells = dist_pred_obs.shape[0]
    if vals.size < 2:
        return np.full(n_cells, float(vals.mean()) if vals.size > 0 else 0.0)
    if np.unique(vals).size == 1:
        return np.full(n_cells, float(vals[0]))
    n = vals.size
    nug, sill, rng, model = params["nugget"], params["sill"], params["range"], params["model"]
    gm = np.full((n + 1, n + 1), float(nug))
    h = dist_obs_obs[np.ix_(avail, avail)]
    off = ~np.eye(n, dtype=bool)
    gm[:n, :n][off] = variogram_models(h, nug, sill, rng, model)[off]
    gm[n, :n] = 1.0
    gm[:n, n] = 1.0
    gm[n, n] = 0.0
    out = np.empty(n_cells)
    d = dist_pred_obs[:, avail]
    for c in range(n_cells):
        gv = np.append(variogram_models(d[c], nug, sill, rng, model), 1.0)
        try:
            w = np.linalg.solve(gm, gv)                       # LAPACK dgesv, like R's solve()
            out[c] = max(0.0, float(np.sum(w[:n] * vals)))
        except np.linalg.LinAlgError:
            out[c] = float(vals.mean())
    return out


def kriging_reference(cx, cy, sx, sy, obs, params):
    """All dates; params[t] as for kriging_one_date.  Returns (n_cells, n_dates), plus the worst 2-norm condition
    number of the kriged dates' gamma matrices (the two solve orders agree to about kappa * 2^-53)."""
    dist = distance_matrix(cx, cy, sx, sy)
    dss = distance_matrix(sx, sy, sx, sy)
    obs = np.asarray(obs, dtype=np.float64)
    out = np.empty((dist.shape[0], obs.shape[0]), dtype=np.float64, order="F")
    for t in range(obs.shape[0]):
        out[:, t] = kriging_one_date(dist, dss, obs[t], params[t])
    return out


def kriging_idw_fallback_reference(cx, cy, sx, sy, obs):
    """All dates of `obs` (n_dates x n_stations) through kriging_idw_fallback_one_date; (n_cells, n_dates)."""
    dist = distance_matrix(cx, cy, sx, sy)
    obs = np.asarray(obs, dtype=np.float64)
    out = np.empty((dist.shape[0], obs.shape[0]), dtype=np.float64, order="F")
    for t in range(obs.shape[0]):
        out[:, t] = kriging_idw_fallback_one_date(dist, obs[t])
    return out


LONG_DOUBLE_ROWSUMS = True


def _row_sums_na_rm(m):
    """rowSums(m, na.rm = TRUE): NaN/NA entries are skipped; R accumulates in long double."""
    if LONG_DOUBLE_ROWSUMS:
        mm = np.where(np.isnan(m), 0.0, m)
        return mm.astype(np.longdouble).sum(axis=1).astype(np.float64)
    return np.nansum(m, axis=1)


# --------------------------------------------------------------------------------------
# Cressman  (R/Cressman.R:216-218, :268-345)
# --------------------------------------------------------------------------------------
def cressman_radii_m(search_radius_km):
    """R/Cressman.R:217-218: km -> m, then sort(unique())."""
    r = np.asarray(search_radius_km, np.float64).reshape(-1) * 1000.0
    return np.unique(r)  # ascending, de-duplicated


def cressman_setup(cx, cy, sx, sy, radii_m):
    """Date-independent part: dist_mat, dist_sq and weights_list (R/Cressman.R:279-300)."""
    cx = np.asarray(cx, np.float64)
    cy = np.asarray(cy, np.float64)
    sx = np.asarray(sx, np.float64)
    sy = np.asarray(sy, np.float64)
    dist_mat = distance_matrix(cx, cy, sx, sy)                      # :279-287
    dist_sq = dist_mat * dist_mat                                   # :288
    weights = []
    for R in radii_m:                                               # :290
        R2 = R * R                                                  # :291
        num = R2 - dist_sq                                          # :293
        den = R2 + dist_sq                                          # :294
        with np.errstate(invalid="ignore", divide="ignore"):
            w = num / den                                           # :295
        w[dist_mat > R] = 0.0                                       # :296
        weights.append(w)
    return weights


def cressman_one_date(weights, obs_vec):
    """crsmn_logic_opt(data_obs) (R/Cressman.R:310-321): one n_cells vector per radius."""
    res = []
    for w in weights:                                               # :312-313
        with np.errstate(invalid="ignore", over="ignore"):
            prod = w * obs_vec[None, :]                             # :314 sweep(w, 2, obs, `*`)
        numer = _row_sums_na_rm(prod)                               # :314 rowSums(na.rm=TRUE)
        denom = _row_sums_na_rm(w)                                  # :315 (ALL stations)
        with np.errstate(invalid="ignore", divide="ignore"):
            o = numer / denom                                       # :316
        o[~np.isfinite(o) | (denom == 0)] = np.nan                  # :317
        res.append(o)
    return res


def cressman_reference(cx, cy, sx, sy, obs, radii_m):
    """Returns a list (one per radius, ascending order as given) of (n_cells, n_dates) arrays."""
    obs = np.asarray(obs, np.float64)
    weights = cressman_setup(cx, cy, sx, sy, radii_m)
    n_cells = weights[0].shape[0]
    outs = [np.empty((n_cells, obs.shape[0]), np.float64) for _ in radii_m]
    for t in range(obs.shape[0]):                                   # :340
        for ri, o in enumerate(cressman_one_date(weights, obs[t])):
            outs[ri][:, t] = o
    return outs


# --------------------------------------------------------------------------------------
# Matrix form of the same two algorithms (Numer = W.Z0, Denom = W.M).  Used where the
# literal per-date loop would take minutes (full-length records in bench.py's parity sample and the
# full-size GPU tests); held to the literal restatement by tests/test_oracle.py.
# --------------------------------------------------------------------------------------
def idw_reference_matrix(cx, cy, sx, sy, obs, p=2.0):
    """(n_cells, n_dates), equal to idw_reference up to the order of summation (BLAS GEMM instead of
    a dgemv per date).  Early-outs R/IDW.R:294-302; cells with an exact zero distance (R/IDW.R:310-326)
    are evaluated by the literal idw_one_date."""
    obs = np.asarray(obs, np.float64)
    dist_mat, w_base = idw_setup(cx, cy, sx, sy, p)
    n_cells, n_dates = dist_mat.shape[0], obs.shape[0]
    present = ~_is_na(obs)                                              # (n_dates, n_st)
    z0 = np.where(present, obs, 0.0)
    zero_rows = np.nonzero((dist_mat == 0).any(axis=1))[0]
    w = w_base.copy()
    w[zero_rows, :] = 0.0                                               # handled below
    with np.errstate(invalid="ignore", divide="ignore", over="ignore"):
        numer = w @ z0.T                                                # :329 / :336 for every date at once
        denom = w @ present.astype(np.float64).T                        # :330 / :337
        out = numer / denom                                             # :331 / :338
    out[~np.isfinite(out) | (denom == 0)] = np.nan                      # :332 / :339
    any_obs = present.any(axis=1)
    all_small = np.array([np.all(np.abs(obs[t][present[t]]) < 1e-12) for t in range(n_dates)])
    out[:, ~any_obs] = np.nan                                           # :299-301
    out[:, any_obs & all_small] = 0.0                                   # :294-297
    if zero_rows.size:
        for t in range(n_dates):
            out[zero_rows, t] = idw_one_date(dist_mat[zero_rows], w_base[zero_rows], obs[t])
    return out


def cressman_reference_matrix(cx, cy, sx, sy, obs, radii_m):
    """List (one per radius) of (n_cells, n_dates): Numer = W_R.Z0 with NA obs skipped (R/Cressman.R:314),
    denominator = row sums of W_R over ALL stations (:315), NA rule :317."""
    obs = np.asarray(obs, np.float64)
    z0 = np.where(_is_na(obs), 0.0, obs)
    outs = []
    for w in cressman_setup(cx, cy, sx, sy, radii_m):
        numer = w @ z0.T
        denom = _row_sums_na_rm(w)[:, None]
        with np.errstate(invalid="ignore", divide="ignore"):
            o = numer / denom
        o[~np.isfinite(o) | (denom == 0)] = np.nan
        outs.append(o)
    return outs


def radius_names(radii_m):
    """R/Cressman.R:357  sprintf("%g km", R/1000)."""
    return ["%g km" % (r / 1000.0) for r in radii_m]


# --------------------------------------------------------------------------------------
# station split  (R/Intr_Connections_module.R:5-52) — manual mode only; the random mode needs
# R's sample() after set.seed(123) (Mersenne-Twister + rejection sampling), which is not restated.
# --------------------------------------------------------------------------------------
def select_manual(station_names, stat_validation):
    """ICM:29,38,42,45: train = names minus stat_validation (order preserved); test = the rest."""
    sv = set([stat_validation] if isinstance(stat_validation, str) else stat_validation)
    train = [n for n in station_names if n not in sv]
    test = [n for n in station_names if n in sv]
    return train, test


# --------------------------------------------------------------------------------------
# R's round(x, digits) for digits > 0 (R >= 4.0.0, src/nmath/fround.c "private_nearbyint" path)
# --------------------------------------------------------------------------------------
def r_round(x, digits=2):
    x = float(x)
    if math.isnan(x) or math.isinf(x):
        return x
    if digits > 15:
        return x
    sgn = 1.0
    if x < 0:
        sgn, x = -1.0, -x
    p10 = 10.0 ** digits
    x10 = p10 * x
    i10 = math.floor(x10)
    xd = i10 / p10
    xu = math.ceil(x10) / p10
    du = xu - x
    dd = x - xd
    if dd < du or (dd == du and (i10 % 2 == 0)):
        return sgn * xd
    return sgn * xu


# --------------------------------------------------------------------------------------
# goodness-of-fit  (R/Intr_gof.R:7-143), metrics() (ICM:55-69), categorical (ICM:73-208)
# --------------------------------------------------------------------------------------
def _complete(sim, obs):
    sim = np.asarray(sim, np.float64)
    obs = np.asarray(obs, np.float64)
    ok = ~(np.isnan(sim) | np.isnan(obs))       # stats::complete.cases
    return sim[ok], obs[ok]


def _sd(v):
    return float(np.std(v, ddof=1))


def _pearson(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    da = a - a.mean()
    db = b - b.mean()
    den = math.sqrt(float((da * da).sum()) * float((db * db).sum()))
    if den == 0:
        return float("nan")     # stats::cor gives NA (with a warning) for zero variance
    return float((da * db).sum() / den)


def _rank_avg(v):
    v = np.asarray(v, np.float64)
    order = np.argsort(v, kind="mergesort")
    ranks = np.empty(len(v), np.float64)
    i = 0
    while i < len(v):
        j = i
        while j + 1 < len(v) and v[order[j + 1]] == v[order[i]]:
            j += 1
        ranks[order[i:j + 1]] = 0.5 * (i + j) + 1.0
        i = j + 1
    return ranks


def gof(sim, obs):
    """The nine metrics of Intr_gof.R, unrounded, as a dict keyed like metrics()' columns."""
    s, o = _complete(sim, obs)
    n = len(o)
    nan = float("nan")
    out = {}
    out["MAE"] = float(np.mean(np.abs(s - o))) if n else nan                  # :7-15
    out["MSE"] = float(np.mean((s - o) ** 2)) if n else nan                   # :18-26
    out["RMSE"] = math.sqrt(out["MSE"]) if n else nan                         # :68-76
    so = float(o.sum()) if n else 0.0
    out["PBIAS %"] = 100.0 * float((s - o).sum()) / so if (n and so != 0) else nan   # :129-143
    if n:
        den = float(((o - o.mean()) ** 2).sum())
        out["NSE"] = 1.0 - float(((s - o) ** 2).sum()) / den if den != 0 else nan    # :110-126
    else:
        out["NSE"] = nan
    if n >= 2:
        ss_tot = float(((o - o.mean()) ** 2).sum())
        out["R2"] = 1.0 - float(((o - s) ** 2).sum()) / ss_tot if ss_tot != 0 else nan  # :51-65
        out["rSpearman"] = _pearson(_rank_avg(o), _rank_avg(s))               # :29-37
        out["rPearson"] = _pearson(o, s)                                      # :40-48
        mo = float(o.mean())
        if mo == 0 or _sd(o) == 0:                                            # :88-96
            out["KGE"] = nan
        else:
            r = _pearson(o, s)
            if math.isnan(r):
                out["KGE"] = nan
            else:
                alpha = _sd(s) / _sd(o)
                beta = float(s.mean()) / mo
                out["KGE"] = 1.0 - math.sqrt((r - 1) ** 2 + (alpha - 1) ** 2 + (beta - 1) ** 2)
    else:
        out["R2"] = out["rSpearman"] = out["rPearson"] = out["KGE"] = nan
    return out


GOF_COLUMNS = ["MAE", "MSE", "RMSE", "PBIAS %", "NSE", "R2", "rSpearman", "rPearson", "KGE"]


def gof_rounded(sim, obs):
    """metrics() gof table row: each metric round(., 2) (ICM:59-69)."""
    g = gof(sim, obs)
    return {k: r_round(g[k], 2) for k in GOF_COLUMNS}


def threshold_categories(rain_threshold):
    """create_threshold_categories (ICM:74-93).  rain_threshold: ordered dict name -> scalar or
    (lo, hi).  Returns (thresholds incl. trailing Inf, category names sorted by lower bound)."""
    names = list(rain_threshold.keys())
    mins = []
    for n in names:
        v = rain_threshold[n]
        mins.append(float(v[0]) if hasattr(v, "__len__") and len(v) == 2 else float(np.asarray(v).reshape(-1)[0]))
    order = sorted(range(len(names)), key=lambda i: mins[i])  # order() is stable
    cats = [names[i] for i in order]
    thr = [mins[i] for i in order] + [float("inf")]
    return thr, cats


def _cut_right_false(v, thr, cats):
    """cut(v, breaks=thr, labels=cats, right=FALSE): [thr[i], thr[i+1]); outside -> NA (None)."""
    out = []
    for x in v:
        if x is None or (isinstance(x, float) and math.isnan(x)):
            out.append(None)
            continue
        lab = None
        for i in range(len(cats)):
            if thr[i] <= x < thr[i + 1]:
                lab = cats[i]
                break
        out.append(lab)
    return out


def categorical_metrics(sim, obs, rain_threshold):
    """calculate_category_metrics (ICM:95-166) for each category, with the operator-precedence
    quirks of HSS (ICM:130-140) and ETS (ICM:143-152) reproduced verbatim.  `sim`/`obs` are the
    merged per-station Obs/Sim columns (NA kept).  data.table `==` on NA_character_ yields NA,
    which is dropped by `[i]` subsetting, so rows with an NA category never count."""
    thr, cats = threshold_categories(rain_threshold)
    oc = _cut_right_false([float(x) for x in obs], thr, cats)
    ec = _cut_right_false([float(x) for x in sim], thr, cats)
    nan = float("nan")
    rows = []
    for cat in cats:
        hits = misses = fa = cn = 0
        for a, b in zip(oc, ec):
            if a is None or b is None:
                continue
            if a == cat and b == cat:
                hits += 1
            elif a == cat and b != cat:
                misses += 1
            elif b == cat and a != cat:
                fa += 1
            else:
                cn += 1
        POD = hits / (hits + misses) if (hits + misses) > 0 else nan
        SR = 1 - (fa / (hits + fa)) if (hits + fa) > 0 else nan
        CSI = hits / (hits + misses + fa) if (hits + misses + fa) > 0 else nan
        HB = (hits + fa) / (hits + misses) if (hits + misses) > 0 else nan
        FAR = fa / (hits + fa) if (hits + fa) > 0 else nan
        HK = POD - (fa / (fa + cn)) if (fa + cn) > 0 else nan       # 0/0 -> NaN in R
        cond = (hits + misses) * (misses + cn) + (hits + fa) * (fa + cn)
        if cond != 0:
            # (2*(h*cn - m*fa)) / (h+m) * (m+cn) + (h+fa)*(fa+cn)   -- precedence as written
            hm = hits + misses
            first = (2 * (hits * cn - misses * fa)) / hm if hm != 0 else _r_div(2 * (hits * cn - misses * fa), 0)
            HSS = first * (misses + cn) + (hits + fa) * (fa + cn)
        else:
            HSS = nan
        tot = hits + misses + fa + cn
        a_random = ((hits + fa) * (hits + misses)) / tot if tot != 0 else nan
        if hits + misses + fa - a_random != 0:                      # NaN != 0 is NA -> ifelse gives NA
            if math.isnan(a_random):
                ETS = nan
            else:
                ETS = hits - _r_div(a_random, hits) + misses + fa - a_random
        else:
            ETS = nan
        rows.append(dict(Category=cat, POD=POD, SR=SR, CSI=CSI, HB=HB, FAR=FAR, HK=HK, HSS=HSS, ETS=ETS))
    return rows


def _r_div(a, b):
    """R's double division (x/0 -> +-Inf, 0/0 -> NaN)."""
    if b == 0:
        if a == 0 or (isinstance(a, float) and math.isnan(a)):
            return float("nan")
        return math.copysign(float("inf"), a)
    return a / b


# --------------------------------------------------------------------------------------
# .validate()  (ICM:215-287): terra::extract at the test stations + per-station metrics
# --------------------------------------------------------------------------------------
def extract_cells(grid, px, py):
    """terra::extract(SpatRaster, points) cell lookup: the cell containing each point, -1 when the
    point lies outside the extent (-> NA row).  Points on the xmax / ymin edge belong to the
    last column / row (terra cellFromXY)."""
    ncol, nrow = grid["ncol"], grid["nrow"]
    xr = (grid["xmax"] - grid["xmin"]) / ncol
    yr = (grid["ymax"] - grid["ymin"]) / nrow
    cells = []
    for x, y in zip(px, py):
        if not (grid["xmin"] <= x <= grid["xmax"] and grid["ymin"] <= y <= grid["ymax"]):
            cells.append(-1)
            continue
        col = int(math.floor((x - grid["xmin"]) / xr))
        row = int(math.floor((grid["ymax"] - y) / yr))
        if x == grid["xmax"]:
            col = ncol - 1
        if y == grid["ymin"]:
            row = nrow - 1
        cells.append(row * ncol + col)
    return cells


def validate_reference(grid, stack, test_x, test_y, test_obs, rain_threshold=None):
    """stack: (n_cells, n_dates) (already n_round-ed, un-masked, R/IDW.R:356-372).
    test_obs: (n_dates, n_test).  Returns list of per-station dicts (gof rounded to 2 digits),
    plus categorical rows when rain_threshold is given.  Duplicated rows created by the
    data.table recycling at ICM:243-246 do not change any metric."""
    cells = extract_cells(grid, test_x, test_y)
    res = []
    for j, c in enumerate(cells):
        sim = np.full(stack.shape[1], np.nan) if c < 0 else stack[c, :]
        row = dict(gof=gof_rounded(sim, test_obs[:, j]))
        if rain_threshold is not None:
            row["categorical"] = categorical_metrics(sim, test_obs[:, j], rain_threshold)
        res.append(row)
    return res
