"""CPU tests of the oracle itself: pinned against the committed golden fixtures, cross-checked
against an independent scalar restatement, and against closed-form cases."""
import math
import os

import numpy as np
import pytest

from conftest import assert_parity
from interpolater_b200 import synthetic
from oracle import reference_port as ref
from oracle import scalar_check


def test_packaged_grid_is_4x6(packaged):
    bbox = packaged["raw"]["bbox"]
    g = ref.grid_from_extent(bbox[0], bbox[1], bbox[2], bbox[3], 5000.0)
    assert (g["ncol"], g["nrow"]) == (4, 6)
    assert g["X"].shape == (24,)
    # row-major from the top-left: first cell is the north-west one
    assert g["X"][0] == pytest.approx(bbox[0] + 2500.0) and g["Y"][0] == pytest.approx(bbox[2] + 6 * 5000.0 - 2500.0)
    assert g["Y"][0] > g["Y"][-1] and g["X"][0] < g["X"][3]


def test_oracle_reproduces_packaged_golden(packaged, packaged_golden):
    z, g = packaged["raw"], packaged_golden
    stations = [str(s) for s in z["stations"]]
    cod = [str(c) for c in z["cod"]]
    xy = {c: (x, y) for c, x, y in zip(cod, z["X"], z["Y"])}
    col = {s: i for i, s in enumerate(stations)}
    train, test = ref.select_manual(stations, "M001")
    tr = sorted(train)
    idw = ref.idw_reference(g["grid_X"], g["grid_Y"], np.array([xy[c][0] for c in tr]), np.array([xy[c][1] for c in tr]),
                            z["obs"][:, [col[c] for c in tr]], 2.0)
    assert_parity(idw, g["idw"], tol=1e-13, label="cfg1 IDW")
    crs = ref.cressman_reference(g["grid_X"], g["grid_Y"], np.array([xy[c][0] for c in train]),
                                 np.array([xy[c][1] for c in train]), z["obs"][:, [col[c] for c in train]], g["radii_m"])
    for i in range(3):
        assert_parity(crs[i], g["cressman"][i], tol=1e-13, label=f"cfg2 Cressman {i}")


def test_packaged_fixture_facts(packaged):
    """60 dates x 10 stations, 8 NA cells (M003:2, M004:2, M007:1, M009:3), max 36.7."""
    obs = packaged["raw"]["obs"]
    assert obs.shape == (60, 10)
    na = np.isnan(obs).sum(axis=0)
    assert na.tolist() == [0, 0, 2, 2, 0, 0, 1, 0, 3, 0]
    assert np.nanmax(obs) == pytest.approx(36.7)


@pytest.mark.skipif(not os.path.exists("/root/reference/data/BD_Obs.rda"), reason="reference not mounted")
def test_rdx3_decoder_matches_frozen_inputs(packaged):
    from oracle import rdata

    df = rdata.data_frame(rdata.load_rda("/root/reference/data/BD_Obs.rda")["BD_Obs"])
    for i, s in enumerate(packaged["stations"]):
        a, b = df[s], packaged["raw"]["obs"][:, i]
        assert np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])
    assert df["Date::class"] == ["Date"] and df["Date"][0] == 16436.0  # 2015-01-01


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_idw_oracle_vs_scalar_restatement(seed):
    cx, cy, sx, sy, obs = synthetic.synthetic_case(7, 5, 9, 10, na_frac=0.2, plant_zero_dist=3, all_na_dates=(2,),
                                                   all_zero_dates=(4,), seed=seed)
    sx[1], sy[1] = sx[0], sy[0]  # two stations co-located on the same cell centre -> running mean
    for p in (2.0, 1.0, 3.5):
        a = ref.idw_reference(cx, cy, sx, sy, obs, p)
        b = np.array(scalar_check.idw_scalar(cx.tolist(), cy.tolist(), sx.tolist(), sy.tolist(), obs.tolist(), p))
        assert_parity(a, b, tol=1e-12, label=f"idw p={p}")
    assert np.all(a[:, 4] == 0.0) and np.all(np.isnan(a[:, 2]))


@pytest.mark.parametrize("seed", [1, 2])
def test_cressman_oracle_vs_scalar_restatement(seed):
    cx, cy, sx, sy, obs = synthetic.synthetic_case(7, 5, 9, 8, na_frac=0.3, all_na_dates=(1,), seed=seed)
    radii = ref.cressman_radii_m([3, 1, 3, 0.4])
    assert radii.tolist() == [400.0, 1000.0, 3000.0]  # sort(unique()), R/Cressman.R:217-218
    a = ref.cressman_reference(cx, cy, sx, sy, obs, radii)
    b = scalar_check.cressman_scalar(cx.tolist(), cy.tolist(), sx.tolist(), sy.tolist(), obs.tolist(), radii.tolist())
    for i in range(len(radii)):
        assert_parity(a[i], np.array(b[i]), tol=1e-12, label=f"cressman {i}")
    # quirk: an all-NA date gives 0 (not NA) wherever a station is in range (R/Cressman.R:314-317)
    in_range = ~np.isnan(a[2][:, 0])
    assert np.all(a[2][in_range, 1] == 0.0)


def test_cressman_unmasked_denominator_pulls_towards_zero():
    """R/Cressman.R:315: the denominator keeps the weight of stations that are NA on the date."""
    cx, cy = np.array([0.0]), np.array([0.0])
    sx, sy = np.array([100.0, -100.0]), np.array([0.0, 0.0])
    obs = np.array([[10.0, 10.0], [10.0, np.nan]])
    out = ref.cressman_reference(cx, cy, sx, sy, obs, np.array([1000.0]))[0]
    assert out[0, 0] == pytest.approx(10.0)
    assert out[0, 1] == pytest.approx(5.0)   # not 10: the NA station still weighs in the denominator


def test_radius_names():
    assert ref.radius_names(np.array([10000.0, 2500.0, 50000.0])) == ["10 km", "2.5 km", "50 km"]


def test_r_round():
    assert ref.r_round(0.125, 2) == 0.12      # tie -> even candidate
    assert ref.r_round(0.135, 2) == 0.14      # 0.135 is above the tie in binary
    assert ref.r_round(2.675, 2) == 2.67      # stored as 2.67499999...
    assert ref.r_round(-1.005, 2) == -1.0
    assert ref.r_round(1.0, 2) == 1.0
    assert math.isnan(ref.r_round(float("nan"), 2))


def test_gof_closed_form():
    obs = np.array([1.0, 2.0, 3.0, 4.0, np.nan])
    sim = np.array([1.5, 2.5, 2.5, 5.0, 9.0])
    g = ref.gof(sim, obs)
    assert g["MAE"] == pytest.approx(0.625)
    assert g["MSE"] == pytest.approx((0.25 * 3 + 1.0) / 4)
    assert g["RMSE"] == pytest.approx(math.sqrt(0.4375))
    assert g["PBIAS %"] == pytest.approx(100 * 1.5 / 10)
    assert g["NSE"] == pytest.approx(1 - 1.75 / 5.0)
    assert g["R2"] == pytest.approx(g["NSE"])
    assert g["rPearson"] == pytest.approx(np.corrcoef(obs[:4], sim[:4])[0, 1])
    assert g["rSpearman"] == pytest.approx(0.9486832980505138)  # ranks (1,2,3,4) vs (1,2.5,2.5,4)
    r = g["rPearson"]
    alpha = np.std(sim[:4], ddof=1) / np.std(obs[:4], ddof=1)
    beta = sim[:4].mean() / 2.5
    assert g["KGE"] == pytest.approx(1 - math.sqrt((r - 1) ** 2 + (alpha - 1) ** 2 + (beta - 1) ** 2))
    empty = ref.gof(np.array([np.nan]), np.array([1.0]))
    assert all(math.isnan(v) for v in empty.values())


def test_extract_cells_edges():
    g = ref.grid_from_extent(0.0, 40.0, 0.0, 30.0, 10.0)
    cells = ref.extract_cells(g, [5.0, 40.0, 0.0, 41.0, 15.0], [25.0, 0.0, 30.0, 5.0, -1.0])
    assert cells == [0, 11, 0, -1, -1]


# ----------------------------------------------------------------------------- oracle properties
# Size-independent properties of the algorithm itself (the GPU suite checks the same ones at full
# size): they guard the restatement against slips that the scalar twin, written from the same
# reading of the reference, could share.
def _case(seed, n_cells=40, n_st=9, n_dates=7, na=0.25):
    rng = np.random.default_rng(seed)
    cx, cy = rng.random(n_cells) * 5e4 + 7e5, rng.random(n_cells) * 5e4 + 9.6e6
    sx, sy = rng.random(n_st) * 5e4 + 7e5, rng.random(n_st) * 5e4 + 9.6e6
    obs = np.round(rng.gamma(0.8, 8.0, (n_dates, n_st)), 1)
    obs[rng.random(obs.shape) < na] = np.nan
    obs[:, 0] = np.where(np.isnan(obs[:, 0]), 1.5, obs[:, 0])       # no all-NA date
    return cx, cy, sx, sy, obs


@pytest.mark.parametrize("seed", range(5))
def test_idw_oracle_properties(seed):
    cx, cy, sx, sy, obs = _case(seed)
    out = ref.idw_reference(cx, cy, sx, sy, obs, 2.0)
    lo, hi = np.nanmin(obs, axis=1), np.nanmax(obs, axis=1)
    assert np.all(out >= lo[None, :] - 1e-9) and np.all(out <= hi[None, :] + 1e-9)   # convex combination
    # a field that is constant over the available stations is reproduced
    const = np.where(np.isnan(obs), np.nan, np.arange(1, obs.shape[0] + 1)[:, None] * 2.5)
    oc = ref.idw_reference(cx, cy, sx, sy, const, 2.0)
    assert np.allclose(oc, (np.arange(1, obs.shape[0] + 1) * 2.5)[None, :], rtol=1e-12)
    # linear in the observations for a fixed availability pattern
    mix = ref.idw_reference(cx, cy, sx, sy, 2.0 * obs + 3.0 * const, 2.0)
    assert np.allclose(mix, 2.0 * out + 3.0 * oc, rtol=1e-11, atol=1e-11)
    # station order only permutes the summation
    perm = np.random.default_rng(seed + 100).permutation(sx.size)
    op = ref.idw_reference(cx, cy, sx[perm], sy[perm], obs[:, perm], 2.0)
    assert np.allclose(op, out, rtol=1e-12)
    # a rigid shift of all coordinates changes nothing beyond rounding of the differences
    os_ = ref.idw_reference(cx + 1024.0, cy - 2048.0, sx + 1024.0, sy - 2048.0, obs, 2.0)
    assert np.allclose(os_, out, rtol=1e-9)
    # p -> large: the nearest available station dominates
    o50 = ref.idw_reference(cx, cy, sx, sy, obs, 60.0)
    d = np.hypot(cx[:, None] - sx[None, :], cy[:, None] - sy[None, :])
    for t in range(obs.shape[0]):
        avail = ~np.isnan(obs[t])
        dn = np.where(avail[None, :], d, np.inf)
        srt = np.sort(dn, axis=1)
        clear = srt[:, 1] > 1.5 * srt[:, 0]        # second-nearest at least 50 % farther: weight ratio < 3e-11
        near = np.argmin(dn, axis=1)
        assert np.allclose(o50[clear, t], obs[t, near[clear]], rtol=1e-9, atol=1e-9)


@pytest.mark.parametrize("seed", range(4))
def test_cressman_oracle_properties(seed):
    cx, cy, sx, sy, obs = _case(seed, na=0.0)
    radii = np.array([6e3, 15e3, 80e3])
    outs = ref.cressman_reference(cx, cy, sx, sy, obs, radii)
    d = np.hypot(cx[:, None] - sx[None, :], cy[:, None] - sy[None, :])
    for R, out in zip(radii, outs):
        none_in_range = ~(d < R).any(axis=1)       # w > 0 only for d < R (d == R gives (R^2-d^2) = 0)
        assert np.array_equal(np.isnan(out[:, 0]), none_in_range)
        assert np.array_equal(np.isnan(out), np.repeat(none_in_range[:, None], obs.shape[0], axis=1))
        ok = ~none_in_range
        assert np.all(out[ok] >= obs.min(axis=1)[None, :].repeat(ok.sum(), 0) - 1e-9)
        assert np.all(out[ok] <= obs.max(axis=1)[None, :].repeat(ok.sum(), 0) + 1e-9)
    # radii are independent passes: one radius alone gives the same stack
    solo = ref.cressman_reference(cx, cy, sx, sy, obs, radii[1:2])[0]
    assert np.array_equal(solo, outs[1], equal_nan=True)
    # an unavailable station still counts in the denominator (R/Cressman.R:315): dropping its value
    # lowers the estimate by exactly its share
    o2 = obs.copy()
    o2[:, 3] = np.nan
    with_na = ref.cressman_reference(cx, cy, sx, sy, o2, radii[2:3])[0]
    w = np.where(d > radii[2], 0.0, (radii[2] ** 2 - d * d) / (radii[2] ** 2 + d * d))
    share = w[:, 3:4] * obs[:, 3][None, :] / w.sum(axis=1)[:, None]
    assert np.allclose(with_na, outs[2] - share, rtol=1e-11, atol=1e-11)


def test_matrix_form_equals_literal_per_date_loop():
    """oracle.idw_reference_matrix / cressman_reference_matrix (used for full-length records) against the
    literal restatement: same NA pattern, same early-out layers, values equal to summation-order noise."""
    from interpolater_b200 import synthetic

    cx, cy, sx, sy, obs = synthetic.synthetic_case(17, 12, 29, 75, na_frac=0.2, plant_zero_dist=3, all_na_dates=(4,),
                                                   all_zero_dates=(9,), seed=5)
    for p in (2.0, 3.5):
        a, b = ref.idw_reference(cx, cy, sx, sy, obs, p), ref.idw_reference_matrix(cx, cy, sx, sy, obs, p)
        assert np.array_equal(np.isnan(a), np.isnan(b))
        m = ~np.isnan(a)
        assert np.allclose(a[m], b[m], rtol=1e-12, atol=0.0)
        assert np.all(b[:, 9] == 0.0) and np.all(np.isnan(b[:, 4]))
    radii = ref.cressman_radii_m([4, 9, 30])
    for a, b in zip(ref.cressman_reference(cx, cy, sx, sy, obs, radii),
                    ref.cressman_reference_matrix(cx, cy, sx, sy, obs, radii)):
        assert np.array_equal(np.isnan(a), np.isnan(b))
        m = ~np.isnan(a)
        assert np.allclose(a[m], b[m], rtol=1e-12, atol=1e-300)


# ---------------------------------------------------------------- Kriging_Ordinary's IDW fallback
def _kriging_case(seed, n_dates=12):
    cx, cy, sx, sy, obs = synthetic.synthetic_case(9, 7, 11, n_dates, na_frac=0.25, plant_zero_dist=2, seed=seed)
    obs[1, :] = np.nan                         # no station available -> 0
    obs[2, :] = np.nan
    obs[2, 4] = 6.5                            # one station -> its value
    obs[3, :] = np.where(np.isnan(obs[3, :]), np.nan, 2.25)   # all available equal -> that value
    return cx, cy, sx, sy, obs


@pytest.mark.parametrize("seed", [3, 4])
def test_kriging_fallback_oracle_rules(seed):
    """R/Kriging_Ordinary.R:406-423 (prelude) and :465-476 (weighted mean, d == 0 -> 1e-10), against a scalar
    per-cell restatement written independently of the vectorised one."""
    cx, cy, sx, sy, obs = _kriging_case(seed)
    got = ref.kriging_idw_fallback_reference(cx, cy, sx, sy, obs)
    assert not np.isnan(got).any()             # the fallback never yields NA
    assert np.all(got[:, 1] == 0.0) and np.all(got[:, 2] == 6.5) and np.all(got[:, 3] == 2.25)
    for t in range(obs.shape[0]):
        v = obs[t]
        ok = ~np.isnan(v)
        if ok.sum() < 2 or np.unique(v[ok]).size == 1:
            continue
        for c in (0, 17, cx.size - 1):
            num = den = 0.0
            for s in np.flatnonzero(ok):
                d = math.sqrt((cx[c] - sx[s]) ** 2 + (cy[c] - sy[s]) ** 2)
                if d == 0:
                    d = 1e-10
                w = 1.0 / (d * d)
                num += w * v[s]
                den += w
            assert abs(got[c, t] - num / den) <= 1e-12 * max(abs(num / den), 1e-12)
    # a cell that coincides with an available station takes (to 1e-20 relative) that station's value,
    # not IDW()'s running mean, and a convex combination stays within the range of the available values
    t = 5
    ok = ~np.isnan(obs[t])
    lo, hi = obs[t][ok].min(), obs[t][ok].max()
    assert np.all(got[:, t] >= lo - 1e-9) and np.all(got[:, t] <= hi + 1e-9)
    for s in (0, 1):
        if ok[s]:
            c = int(np.flatnonzero((cx == sx[s]) & (cy == sy[s]))[0])
            assert abs(got[c, t] - obs[t, s]) <= 1e-12 * max(abs(obs[t, s]), 1.0)


def test_kriging_oracle_literal_solve_equals_one_solve_per_date():
    """The kriging branch of perform_kriging (R/Kriging_Ordinary.R:428-497), literal per-cell solve, against the
    identity the engine uses — sum(w v) = alpha' [gamma(c); 1] with alpha = Gamma^-1 [v; 0] — and against the
    interpolation property of kriging with a zero nugget effect on the diagonal... which the reference does NOT
    have: its diagonal holds the nugget, so a cell on a station does not reproduce the observation exactly."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(8, 6, 9, 5, na_frac=0.2, plant_zero_dist=1, seed=11)
    obs[0, :] = np.where(np.isnan(obs[0, :]), np.nan, 3.0)            # constant date
    params = [dict(nugget=0.4, sill=30.0, range=9000.0, model=m) for m in ("exponential", "spherical", "gaussian", "linear")]
    params.append(dict(use_idw=True))
    got = ref.kriging_reference(cx, cy, sx, sy, obs, params)
    assert np.all(got[:, 0] == 3.0) and not np.isnan(got).any() and np.all(got >= 0.0)
    assert_parity(got[:, 4], ref.kriging_idw_fallback_one_date(ref.distance_matrix(cx, cy, sx, sy), obs[4]), label="use_idw")
    dist = ref.distance_matrix(cx, cy, sx, sy)
    dss = ref.distance_matrix(sx, sy, sx, sy)
    for t in (1, 2, 3):
        ok = ~np.isnan(obs[t])
        v = obs[t][ok]
        if v.size < 2 or np.unique(v).size == 1:
            continue
        n = v.size
        p = params[t]
        gm = np.full((n + 1, n + 1), p["nugget"])
        h = dss[np.ix_(ok, ok)]
        off = ~np.eye(n, dtype=bool)
        gm[:n, :n][off] = ref.variogram_models(h, p["nugget"], p["sill"], p["range"], p["model"])[off]
        gm[n, :n] = gm[:n, n] = 1.0
        gm[n, n] = 0.0
        alpha = np.linalg.solve(gm, np.append(v, 0.0))
        g = ref.variogram_models(dist[:, ok], p["nugget"], p["sill"], p["range"], p["model"])
        want = np.maximum(0.0, g @ alpha[:n] + alpha[n])
        kappa = np.linalg.cond(gm)
        assert np.allclose(got[:, t], want, rtol=max(1e-10, 50 * kappa * 2.0 ** -53), atol=1e-12)
