"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the oracle on
the same seeded inputs, against the committed golden fixtures, and — at BASELINE.json's full size —
through size-independent properties.  Tolerance (SURVEY §8c / north_star): 1e-9 relative in double,
NA positions exact, exact-zero layers bit-exact."""
import os

import numpy as np
import pytest

from conftest import RAIN_THRESHOLD, assert_parity
import interpolater_b200 as ipr
from interpolater_b200 import _native, engine, synthetic
from oracle import reference_port as ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def _require_device():
    lib = _native.load()
    assert lib.ipr_device_count() >= 1, "no sm_100 device: the engine has no CPU fallback"


# ---------------------------------------------------------------- seeded parity vs the oracle
@pytest.mark.parametrize("shape", [(64, 64, 64, 32), (5, 3, 1, 1), (1, 1, 3, 7), (33, 7, 17, 130), (30, 11, 37, 700)])
def test_idw_matches_oracle(shape):
    ncol, nrow, n_st, n_dates = shape
    cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, n_st, n_dates, na_frac=0.1 if n_st > 1 else 0.0,
                                                   plant_zero_dist=min(3, n_st, ncol * nrow),
                                                   all_na_dates=(3,) if n_dates > 3 else (),
                                                   all_zero_dates=(5,) if n_dates > 5 else (), seed=shape[2])
    got = engine.idw_grid(cx, cy, sx, sy, obs, p=2.0)
    assert_parity(got, ref.idw_reference(cx, cy, sx, sy, obs, 2.0), label=f"idw {shape}")
    if n_dates > 5:
        assert np.all(got[:, 5] == 0.0), "all-zero date must give a bit-exact zero layer"
        assert np.all(np.isnan(got[:, 3])), "all-NA date must give an NA layer"


def test_idw_output_na_is_r_na_real():
    cx, cy, sx, sy, obs = synthetic.synthetic_case(4, 4, 5, 3, all_na_dates=(1,))
    got = engine.idw_grid(cx, cy, sx, sy, obs)
    bits = got[:, 1].view(np.uint64)
    assert np.all(bits == 0x7FF00000000007A2)


@pytest.mark.parametrize("p", [1.0, 1.5, 2.0, 3.0, 4.0, 6.0])
def test_idw_powers(p):
    cx, cy, sx, sy, obs = synthetic.synthetic_case(20, 13, 23, 40, na_frac=0.15, plant_zero_dist=2, seed=int(p * 10))
    assert_parity(engine.idw_grid(cx, cy, sx, sy, obs, p=p), ref.idw_reference(cx, cy, sx, sy, obs, p), label=f"p={p}")


def test_idw_tile_variants_mixed():
    """NA-free 256/128-date tiles and masked 128-date tiles in one call."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(30, 11, 37, 900, plant_zero_dist=2, seed=5)
    obs[130:140, 3] = np.nan
    obs[600, :] = np.nan
    obs[899, 0] = np.nan
    assert_parity(engine.idw_grid(cx, cy, sx, sy, obs), ref.idw_reference(cx, cy, sx, sy, obs, 2.0), label="mixed tiles")


def test_idw_colocated_stations_running_mean():
    """Several stations on one cell centre, some NA on some dates (R/IDW.R:310-326)."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(8, 8, 12, 20, na_frac=0.3, seed=11)
    sx[:3], sy[:3] = cx[10], cy[10]
    sx[3], sy[3] = cx[40], cy[40]
    obs[4, :3] = np.nan           # all co-located stations missing -> plain IDW for that cell/date
    got = engine.idw_grid(cx, cy, sx, sy, obs)
    assert_parity(got, ref.idw_reference(cx, cy, sx, sy, obs, 2.0), label="co-located")


def test_idw_large_utm_coordinates_and_tiny_distances():
    rng = np.random.default_rng(3)
    cx, cy = synthetic.regular_grid(16, 16, res_m=25.0, xmin=706058.23, ymin=9680563.48)
    sx = 706058.23 + rng.random(40) * 400
    sy = 9680563.48 + rng.random(40) * 400
    sx[0], sy[0] = cx[5] + 1e-6, cy[5]      # a micrometre away: huge but finite weight
    obs = np.round(rng.gamma(0.8, 8.0, (12, 40)), 1)
    assert_parity(engine.idw_grid(cx, cy, sx, sy, obs), ref.idw_reference(cx, cy, sx, sy, obs, 2.0), label="utm")


def test_idw_signed_data_conditioned_tolerance():
    """Signed observations cancel: judged on sum w|z| / sum w as SURVEY §8c prescribes."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(12, 12, 30, 16, seed=21)
    rng = np.random.default_rng(21)
    obs = obs * rng.choice([-1.0, 1.0], size=obs.shape)
    got = engine.idw_grid(cx, cy, sx, sy, obs)
    want = ref.idw_reference(cx, cy, sx, sy, obs, 2.0)
    scale = ref.idw_reference(cx, cy, sx, sy, np.abs(obs), 2.0)
    assert np.all(np.abs(got - want) <= 1e-9 * np.maximum(scale, 1e-12))


@pytest.mark.parametrize("radii_km", [[10], [20, 10], [50, 20, 10], [3, 1.5, 40, 3]])
def test_cressman_matches_oracle(radii_km):
    cx, cy, sx, sy, obs = synthetic.synthetic_case(30, 20, 41, 150, na_frac=0.1, all_na_dates=(4,), seed=len(radii_km))
    radii = ref.cressman_radii_m(radii_km)
    got = engine.cressman_grid(cx, cy, sx, sy, obs, radii)
    want = ref.cressman_reference(cx, cy, sx, sy, obs, radii)
    assert got.shape == (len(radii), cx.size, 150)
    for i in range(len(radii)):
        assert_parity(got[i], want[i], label=f"cressman {radii[i]}")
    # all-NA date: 0 where any station is in range, NA elsewhere (unmasked denominator quirk)
    big = got[-1][:, 4]
    assert np.all((big == 0.0) | np.isnan(big))


@pytest.mark.parametrize("xdiv", ["1", "4", "16", "64"])
def test_cressman_every_cell_tile_shape_matches_oracle(monkeypatch, xdiv):
    """The four cell-tile shapes of the fused kernel (8 x 8, 4 x 16, 2 x 32, 1 x 64 cells; chosen per radius in
    production) forced in turn, on a grid whose width is no multiple of a run and with long and short lists,
    duplicate-free ragged date range (NP = 6 tail) included; and the round-1 path behind IPR_CRESSMAN_LEGACY."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(83, 37, 150, 330, na_frac=0.07, all_na_dates=(9,), seed=int(xdiv))
    radii = ref.cressman_radii_m([2, 9, 35, 120])
    want = ref.cressman_reference(cx, cy, sx, sy, obs, radii)
    monkeypatch.setenv("IPR_CR_XDIV", xdiv)
    got = engine.cressman_grid(cx, cy, sx, sy, obs, radii)
    for i in range(len(radii)):
        assert_parity(got[i], want[i], label=f"cressman xdiv={xdiv} R={radii[i]:.0f}")


@pytest.mark.parametrize("xdiv", ["1", "4", "16", "64"])
@pytest.mark.parametrize("grid", [(40, 27), (64, 16), (8, 9)])
def test_cressman_rectangular_tiles_on_rasters_with_ragged_edges(monkeypatch, xdiv, grid):
    """Rasters whose rows are whole numbers of 8-cell runs get exact rectangular tiles (regular_run_perm): 5, 8
    or 1 runs per row against tile widths of 1 / 2 / 4 / 8 runs and row counts that are no multiple of the tile
    height, so that full tiles, ragged right / bottom edges and the padding runs all occur."""
    ncol, nrow = grid
    cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, 90, 140, na_frac=0.05, seed=ncol + int(xdiv))
    radii = ref.cressman_radii_m([3, 12, 60])
    want = ref.cressman_reference(cx, cy, sx, sy, obs, radii)
    monkeypatch.setenv("IPR_CR_XDIV", xdiv)
    got = engine.cressman_grid(cx, cy, sx, sy, obs, radii)
    for i in range(len(radii)):
        assert_parity(got[i], want[i], label=f"raster {grid} xdiv={xdiv} R={radii[i]:.0f}")
    monkeypatch.setenv("IPR_CR_MORTON", "1")          # the general (Morton) tiles on the same raster
    got_m = engine.cressman_grid(cx, cy, sx, sy, obs, radii)
    for i in range(len(radii)):
        assert_parity(got_m[i], want[i], label=f"raster {grid} Morton xdiv={xdiv} R={radii[i]:.0f}")


def test_cressman_weights_at_the_cut_off_are_the_references():
    """A station exactly R away from a cell centre (d == R: weight 0/(2 R^2) = 0, still 'in range' for the
    denominator test) and stations a rounding error inside / outside the radius: the fused kernel computes
    dist = sqrt(dx^2 + dy^2), dist_sq = dist * dist and the cut-off dist > R like R/Cressman.R:279-300."""
    cx, cy = synthetic.regular_grid(12, 9)
    R = 5000.0
    sx = np.array([cx[0] + R, cx[5] + 3000.0, cx[17] + R * (1 + 2e-16), cx[30] - R * (1 - 2e-16), cx[40] + 1.0])
    sy = np.array([cy[0], cy[5] + 4000.0, cy[17], cy[30], cy[40] + 1.0])
    rng = np.random.default_rng(4)
    obs = np.round(rng.gamma(0.8, 8.0, (40, sx.size)), 1) + 0.1
    got = engine.cressman_grid(cx, cy, sx, sy, obs, np.array([R]))
    assert_parity(got[0], ref.cressman_reference(cx, cy, sx, sy, obs, np.array([R]))[0], label="cut-off")


def test_cressman_radius_without_neighbours_is_all_na():
    cx, cy, sx, sy, obs = synthetic.synthetic_case(6, 6, 4, 5, seed=2)
    got = engine.cressman_grid(cx, cy, sx + 1e6, sy, obs, np.array([1000.0]))
    assert np.isnan(got).all()


@pytest.mark.parametrize("seed", range(24))
def test_randomised_shapes_against_oracle(seed):
    """Ragged sizes around every tiling boundary (64-cell tiles, 16-station stages, 64/128-date
    tiles), random NA density, co-located stations, random power / radii."""
    rng = np.random.default_rng(1000 + seed)
    ncol, nrow = int(rng.integers(1, 24)), int(rng.integers(1, 14))
    n_st = int(rng.choice([1, 2, 15, 16, 17, 31, 33, 48, 70]))
    n_dates = int(rng.choice([1, 2, 31, 32, 33, 63, 64, 65, 96, 127, 128, 129, 191, 193, 257, 300]))
    na_frac = float(rng.choice([0.0, 0.0, 0.03, 0.5]))
    cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, n_st, n_dates, na_frac=na_frac,
                                                   plant_zero_dist=int(min(n_st, ncol * nrow, rng.integers(0, 3))),
                                                   seed=5000 + seed)
    if n_dates > 40 and na_frac == 0.0 and seed % 2:
        obs[int(rng.integers(0, n_dates)), int(rng.integers(0, n_st))] = np.nan   # a single NA -> one masked granule
    p = float(rng.choice([1.0, 2.0, 2.0, 2.5, 4.0]))
    assert_parity(engine.idw_grid(cx, cy, sx, sy, obs, p=p), ref.idw_reference(cx, cy, sx, sy, obs, p),
                  label=f"idw seed={seed} {ncol}x{nrow} S={n_st} D={n_dates} na={na_frac} p={p}")
    radii = ref.cressman_radii_m(rng.choice([0.7, 1.5, 3, 6, 12, 40], size=int(rng.integers(1, 4)), replace=False))
    got = engine.cressman_grid(cx, cy, sx, sy, obs, radii)
    want = ref.cressman_reference(cx, cy, sx, sy, obs, radii)
    for i in range(len(radii)):
        assert_parity(got[i], want[i], label=f"cressman seed={seed} R={radii[i]}")


def test_long_record_on_a_small_grid():
    """Many dates (several launches of <= 56 date tiles each), few cells and stations — the shape of
    the packaged fixture stretched in time; mixes NA-free, sparse-NA and NA-heavy granules."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(6, 4, 9, 9000, seed=41)
    rng = np.random.default_rng(41)
    obs[2000:2400][rng.random((400, 9)) < 0.3] = np.nan      # NA-heavy stretch -> masked GEMM
    obs[5000:5010, 2] = np.nan                                # a few NA -> corrected denominators
    obs[7777, :] = np.nan
    assert_parity(engine.idw_grid(cx, cy, sx, sy, obs), ref.idw_reference(cx, cy, sx, sy, obs, 2.0), label="D=9000")
    radii = np.array([1500.0, 4000.0])
    got = engine.cressman_grid(cx, cy, sx, sy, obs, radii)
    want = ref.cressman_reference(cx, cy, sx, sy, obs, radii)
    for i in range(2):
        assert_parity(got[i], want[i], label=f"D=9000 cressman {i}")


def test_many_stations_and_wide_station_extent():
    """More stations than one shared-memory chunk of the weight kernel (1024) and stations far
    outside the grid (Morton order must cope with a bounding box much larger than the raster)."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(9, 7, 1100, 5, na_frac=0.1, seed=31)
    sx[::7] += 3.0e6
    sy[::11] -= 2.0e6
    assert_parity(engine.idw_grid(cx, cy, sx, sy, obs), ref.idw_reference(cx, cy, sx, sy, obs, 2.0), label="S=1100")
    radii = np.array([2000.0, 9000.0])
    got = engine.cressman_grid(cx, cy, sx, sy, obs, radii)
    want = ref.cressman_reference(cx, cy, sx, sy, obs, radii)
    for i in range(2):
        assert_parity(got[i], want[i], label=f"S=1100 cressman {i}")


def test_cressman_long_station_lists():
    """The branches of the fused Cressman kernel that the other cases never reach: a tile whose station list is
    longer than its shared-memory copy (2,048 entries: the refilling warp reads the list from global memory),
    more listed stations than one coordinate chunk of the panel generation (2,048), and more 8-station groups
    than one candidate round holds (4,096 groups = 32,768 stations)."""
    rng = np.random.default_rng(77)
    cx, cy = synthetic.regular_grid(11, 7)
    n_st = 33500                                   # 4,188 groups: two candidate rounds
    sx = cx.min() + rng.random(n_st) * 40e3 - 15e3
    sy = cy.min() + rng.random(n_st) * 40e3 - 15e3
    obs = np.round(rng.gamma(0.8, 8.0, (6, n_st)), 1)
    obs[rng.random(obs.shape) < 0.1] = np.nan
    radii = np.array([1500.0, 60e3])               # ~100 stations in range / all 33,500 of them
    got = engine.cressman_grid(cx, cy, sx, sy, obs, radii)
    want = ref.cressman_reference(cx, cy, sx, sy, obs, radii)
    for i in range(2):
        assert_parity(got[i], want[i], label=f"S=33500 cressman R={radii[i]:.0f}")


def test_duplicate_cells_and_all_stations_on_cells():
    """Every station sits on a cell centre and some cells are listed twice (the C ABI takes arbitrary
    cell arrays): zero-distance bookkeeping with many pairs."""
    cx, cy = synthetic.regular_grid(6, 5)
    cx = np.concatenate([cx, cx[:7]])
    cy = np.concatenate([cy, cy[:7]])
    rng = np.random.default_rng(8)
    idx = rng.choice(30, 12, replace=False)
    sx, sy = cx[idx].copy(), cy[idx].copy()
    obs = np.round(rng.gamma(0.8, 8.0, (9, 12)), 1)
    obs[rng.random(obs.shape) < 0.25] = np.nan
    assert_parity(engine.idw_grid(cx, cy, sx, sy, obs), ref.idw_reference(cx, cy, sx, sy, obs, 2.0), label="dups")


def test_device_post_pass_rounding_and_mask():
    """Rounding (R's round / terra's round) and the polygon mask fused on the device equal the host
    definitions applied to the un-rounded engine output."""
    from interpolater_b200 import spatial

    cx, cy, sx, sy, obs = synthetic.synthetic_case(37, 23, 40, 150, na_frac=0.1, all_na_dates=(4,), seed=3)
    raw = engine.idw_grid(cx, cy, sx, sy, obs)
    for digits in (0, 1, 2, 3):
        got = engine.idw_grid(cx, cy, sx, sy, obs, n_round=digits, round_mode=engine.ROUND_R)
        want = np.array([ref.r_round(v, digits) for v in raw.ravel()]).reshape(raw.shape)
        assert np.array_equal(got, want, equal_nan=True), f"R round digits={digits}"
        got_t = engine.idw_grid(cx, cy, sx, sy, obs, n_round=digits, round_mode=engine.ROUND_TERRA)
        assert np.array_equal(got_t, spatial.terra_round(raw, digits), equal_nan=True), f"terra round {digits}"
    assert np.array_equal(spatial.r_round(raw, 2), engine.idw_grid(cx, cy, sx, sy, obs, n_round=2), equal_nan=True)
    keep = (np.arange(cx.size) % 3 != 0).astype(np.uint8)
    got = engine.idw_grid(cx, cy, sx, sy, obs, n_round=1, cell_keep=keep)
    want = spatial.r_round(raw, 1)
    want[keep == 0, :] = np.nan
    assert np.array_equal(got, want, equal_nan=True)
    assert np.all(got[keep == 0, :].view(np.uint64) == 0x7FF00000000007A2)
    radii = np.array([6e3, 15e3])
    rawc = engine.cressman_grid(cx, cy, sx, sy, obs, radii)
    gotc = engine.cressman_grid(cx, cy, sx, sy, obs, radii, n_round=2, cell_keep=keep)
    wantc = spatial.terra_round(np.ascontiguousarray(rawc), 2)
    wantc[:, keep == 0, :] = np.nan
    assert np.array_equal(np.ascontiguousarray(gotc), wantc, equal_nan=True)


def test_idw_mask_equals_crop_mask_of_unmasked_run(packaged):
    """IDW(Mask=TRUE) computes only the cropped window and masks on the device; it must equal
    terra::mask(terra::crop(x, shp, snap="out"), shp) applied to the Mask=FALSE stack."""
    args = dict(BD_Obs=packaged["BD_Obs"], BD_Coord=packaged["BD_Coord"], shapefile=packaged["shapefile"],
                grid_resolution=2, p=2, n_round=1)
    full = ipr.IDW(Mask=False, **args)
    masked = ipr.IDW(Mask=True, **args)
    want = full.crop_mask(packaged["shapefile"])
    assert (masked.geom.ncol, masked.geom.nrow) == (want.geom.ncol, want.geom.nrow)
    assert masked.geom.xmin == want.geom.xmin and masked.geom.ymax == want.geom.ymax
    assert np.array_equal(masked.values, want.values, equal_nan=True)
    assert np.isnan(masked.values).any() and not np.isnan(masked.values).all()


def test_sparse_na_correction_matches_masked_gemm_and_oracle(monkeypatch):
    """Granules with few NA are contracted unmasked and their denominators corrected; the result must
    agree with the masked-GEMM path and with the oracle, including when a missing station carries
    almost all of a cell's weight (the correction then sums the available weights directly)."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(33, 21, 90, 300, na_frac=0.02, plant_zero_dist=2, seed=17)
    sx[5], sy[5] = cx[100] + 1e-3, cy[100]          # a millimetre from a cell centre: w ~ 1e6
    obs[10:40, 5] = np.nan                           # ... and missing on 30 dates
    obs[200, :] = np.nan                             # an all-NA date inside a sparse granule
    obs[0:2, 0] = np.nan                             # co-located (planted) station missing
    want = ref.idw_reference(cx, cy, sx, sy, obs, 2.0)
    monkeypatch.setenv("IPR_SPARSE_NA", "1")
    plan = engine.Plan(0, cx, cy, sx, sy, 300)
    plan.set_obs(obs)
    l0 = plan.launch_count
    import torch
    o = torch.empty((300, cx.size), dtype=torch.float64, device="cuda")
    plan.idw(o.data_ptr(), 2.0)
    plan.sync()
    sparse = o.cpu().numpy().T
    plan.close()
    monkeypatch.setenv("IPR_SPARSE_NA", "0")
    dense = engine.idw_grid(cx, cy, sx, sy, obs)
    assert_parity(sparse, want, label="sparse-NA path")
    assert_parity(dense, want, label="masked-GEMM path")
    assert_parity(sparse, dense, tol=1e-12, label="sparse vs masked")
    for p in (3.0, 4.0):
        monkeypatch.setenv("IPR_SPARSE_NA", "1")
        a = engine.idw_grid(cx, cy, sx, sy, obs, p=p)
        assert_parity(a, ref.idw_reference(cx, cy, sx, sy, obs, p), label=f"sparse p={p}")


def test_int8_mask_contraction_matches_fp64_and_oracle(monkeypatch):
    """Default path (IPR_MASK_INT8=0 disables): the availability-mask GEMM runs error-free on the tcgen05 integer tensor path (8-bit
    digit planes of the weights); same 1e-9 bar against the oracle, and agreement with the FP64
    masked GEMM far below it, including a cell whose dominant station is missing (direct re-evaluation)."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(29, 23, 75, 300, na_frac=0.25, plant_zero_dist=2,
                                                   all_na_dates=(17,), all_zero_dates=(33,), seed=23)
    sx[7], sy[7] = cx[200] + 1e-4, cy[200]        # 0.1 mm from a cell centre: w ~ 1e8 against ~1e-7 for the rest
    obs[50:120, 7] = np.nan
    want = ref.idw_reference(cx, cy, sx, sy, obs, 2.0)
    monkeypatch.setenv("IPR_MASK_INT8", "0")
    fp64 = engine.idw_grid(cx, cy, sx, sy, obs)
    monkeypatch.setenv("IPR_MASK_INT8", "1")
    int8 = engine.idw_grid(cx, cy, sx, sy, obs)
    assert_parity(fp64, want, label="fp64 masked")
    assert_parity(int8, want, label="int8 masked")
    assert_parity(int8, fp64, tol=1e-11, label="int8 vs fp64 masked")
    assert np.all(int8[:, 33] == 0.0) and np.all(np.isnan(int8[:, 17]))
    for p in (1.0, 4.0):
        assert_parity(engine.idw_grid(cx, cy, sx, sy, obs, p=p), ref.idw_reference(cx, cy, sx, sy, obs, p),
                      label=f"int8 masked p={p}")
    # ragged windows through the plan API
    import torch
    plan = engine.Plan(0, cx, cy, sx, sy, 300)
    plan.set_obs(obs)
    o = torch.full((300, cx.size), -1.0, dtype=torch.float64, device="cuda")
    for lo, hi in [(0, 70), (70, 257), (257, 300)]:
        plan.idw(o[lo:].data_ptr(), 2.0, t0=lo, t1=hi)
    plan.sync()
    assert_parity(o.cpu().numpy().T, want, label="int8 masked, windows")
    plan.close()


def test_int8_mask_list_overflow_falls_back_to_fp64_gemm(monkeypatch):
    """Peaked weights (p = 6) with many NA: far more (cell, date) pairs fall below 2^-10 of their row scale than
    the direct-fix list holds (capacity forced to 8 here).  The launch group must switch to the FP64 masked GEMM
    on the device by itself — same parity bar, no failed call (round-1 behaviour: IPR_EUNSUPPORTED)."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(40, 33, 60, 260, na_frac=0.3, plant_zero_dist=1, seed=77)
    want = ref.idw_reference(cx, cy, sx, sy, obs, 6.0)
    monkeypatch.setenv("IPR_FB_CAP", "8")
    got = engine.idw_grid(cx, cy, sx, sy, obs, p=6.0)
    assert_parity(got, want, label="int8 masked, overflowing fix list")
    monkeypatch.delenv("IPR_FB_CAP")
    assert_parity(engine.idw_grid(cx, cy, sx, sy, obs, p=6.0), want, label="int8 masked, roomy fix list")


@pytest.mark.parametrize("p", [0.05, 0.1, 0.5])
def test_idw_small_powers_padding_carries_no_weight(p):
    """p < 1: a padded station at (1e150, 1e150) would weigh pow(2e300, -p/2) ~ 3e-8 (p = 0.05) if it were
    left to underflow; station counts that are not multiples of 16 exercise the padding, NA-free, sparse-NA
    and masked granules alike."""
    for n_st, na in [(23, 0.0), (37, 0.01), (50, 0.3)]:
        cx, cy, sx, sy, obs = synthetic.synthetic_case(19, 14, n_st, 140, na_frac=na, seed=n_st)
        assert_parity(engine.idw_grid(cx, cy, sx, sy, obs, p=p), ref.idw_reference(cx, cy, sx, sy, obs, p),
                      label=f"p={p} n_st={n_st} na={na}")


def test_buffer_pool_release():
    lib = _native.load()
    cx, cy, sx, sy, obs = synthetic.synthetic_case(300, 300, 64, 130)
    a = engine.idw_grid(cx, cy, sx, sy, obs)
    lib.ipr_release_cached_memory()
    b = engine.idw_grid(cx, cy, sx, sy, obs)
    assert np.array_equal(a, b, equal_nan=True)


# ---------------------------------------------------------------- golden fixtures
def test_synthetic_golden(synthetic_golden):
    g = synthetic_golden
    cx, cy, sx, sy, ob = synthetic.synthetic_case(24, 20, 40, 24, na_frac=0.1, plant_zero_dist=4, all_na_dates=(3,),
                                                  all_zero_dates=(7,))
    assert_parity(engine.idw_grid(cx, cy, sx, sy, ob, 2.0), g["a_idw"], label="a_idw")
    assert_parity(engine.idw_grid(cx, cy, sx, sy, ob, 3.0), g["a_idw_p3"], label="a_idw_p3")
    got = engine.cressman_grid(cx, cy, sx, sy, ob, ref.cressman_radii_m([2, 4, 8, 30]))
    for i in range(4):
        assert_parity(got[i], g["a_crs"][i], label=f"a_crs[{i}]")
    cx, cy, sx, sy, ob = synthetic.synthetic_case(20, 9, 37, 280, plant_zero_dist=2)
    assert_parity(engine.idw_grid(cx, cy, sx, sy, ob, 2.0), g["b_idw"], label="b_idw")


def test_packaged_idw_cfg1(packaged, packaged_golden, capfd):
    """BASELINE.json configs[0]: IDW on the packaged data, grid_resolution=5, p=2,
    stat_validation='M001' — through the drop-in signature."""
    out = ipr.IDW(packaged["BD_Obs"], packaged["BD_Coord"], packaged["shapefile"], grid_resolution=5, p=2,
                  Mask=False, training=1, stat_validation="M001")
    ens, val = out["Ensamble"], out["Validation"]
    assert ens.nlyr() == 60 and ens.names[0] == "2015-01-01" and ens.names[-1] == "2015-03-01"
    assert_parity(ens.values, packaged_golden["idw"], label="cfg1 IDW")
    assert val["ID"].tolist() == ["M001"]
    assert np.array_equal(val[ref.GOF_COLUMNS].to_numpy()[0], packaged_golden["idw_gof"][0])
    err = capfd.readouterr().err
    assert "Analysis in progress. Please wait..." in err and "Validation process in progress" in err


def test_packaged_idw_reference_test_expectations(packaged, tmp_path, capfd, monkeypatch):
    """tests/testthat/test-IDW.R: SpatRaster with 60 layers with/without validation; n_round; mask;
    save_model writes <name>.nc / Model_IDW.nc and says "Model saved successfully"."""
    monkeypatch.chdir(tmp_path)
    out = ipr.IDW(packaged["BD_Obs"], packaged["BD_Coord"], packaged["shapefile"], grid_resolution=5, p=2, n_round=1,
                  training=1, stat_validation="M004", Rain_threshold=RAIN_THRESHOLD)
    assert isinstance(out["Ensamble"], ipr.SpatRaster) and out["Ensamble"].nlyr() == 60
    assert set(out["Validation"]) == {"gof", "categorical_metrics"}
    ens = ipr.IDW(packaged["BD_Obs"], packaged["BD_Coord"], packaged["shapefile"], grid_resolution=5, p=2, n_round=1,
                  save_model=True)
    assert isinstance(ens, ipr.SpatRaster) and ens.nlyr() == 60
    v = ens.values[~np.isnan(ens.values)]
    assert np.allclose(v * 10, np.round(v * 10), atol=1e-9)          # n_round = 1
    assert np.isnan(ens.values).any()                                 # masked by the polygon
    assert os.path.exists("Model_IDW.nc")
    ipr.IDW(packaged["BD_Obs"], packaged["BD_Coord"], packaged["shapefile"], grid_resolution=5, save_model=True,
            name_save="custom")
    assert os.path.exists("custom.nc")
    err = capfd.readouterr().err
    assert "Model saved successfully" in err and "Training not specified; using all data for training." in err


def test_packaged_cressman_cfg2(packaged, packaged_golden, tmp_path, monkeypatch, capfd):
    """BASELINE.json configs[1]: Cressman on the packaged data, search_radius=c(50,20,10)."""
    monkeypatch.chdir(tmp_path)
    out = ipr.Cressman(packaged["BD_Obs"], packaged["BD_Coord"], packaged["shapefile"], grid_resolution=5,
                       search_radius=[50, 20, 10], stat_validation="M001", Rain_threshold=RAIN_THRESHOLD,
                       save_model=True)
    ens, val = out["Ensamble"], out["Validation"]
    assert list(ens) == ["10 km", "20 km", "50 km"] == list(val)
    for i, nm in enumerate(ens):
        assert ens[nm].nlyr() == 60
        assert_parity(ens[nm].values, packaged_golden["cressman"][i], label=f"cfg2 {nm}")
        assert np.array_equal(val[nm]["gof"][ref.GOF_COLUMNS].to_numpy()[0], packaged_golden["cressman_gof"][i, 0])
        assert os.path.exists(f"Radius_{nm}.nc")
    assert "Model saved successfully" in capfd.readouterr().err
    # tests/testthat/test-Cressman.R:38-59: random split + n_round
    out = ipr.Cressman(packaged["BD_Obs"], packaged["BD_Coord"], packaged["shapefile"], grid_resolution=5,
                       search_radius=10, training=0.8, n_round=2)
    assert out["Ensamble"]["10 km"].nlyr() == 60 and out["Validation"]["10 km"]["ID"].tolist() == ["M004", "M005"]


def test_synthetic_masked_holdout_validation_cfg5_twin():
    """Small-scale twin of BASELINE.json configs[4]: 20 % per-date NA, 20 % of the stations held out
    and passed as stat_validation; raster and validation table through the drop-in signature."""
    import pandas as pd

    ncol, nrow, n_st, n_dates = 40, 30, 60, 140
    cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, n_st, n_dates, na_frac=0.2, seed=77)
    names = [f"S{i:03d}" for i in range(n_st)]
    dates = pd.date_range("2001-01-01", periods=n_dates, freq="D")
    BD_Obs = pd.DataFrame({"Date": dates, **{nm: obs[:, i] for i, nm in enumerate(names)}})
    BD_Coord = pd.DataFrame({"Cod": names, "X": sx, "Y": sy, "Z": 0.0})
    held = [names[i] for i in np.random.default_rng(20260104).choice(n_st, n_st // 5, replace=False)]
    shp = ipr.SpatVector.from_bbox(200000.0, 200000.0 + ncol * 1000.0, 9000000.0, 9000000.0 + nrow * 1000.0)
    out = ipr.IDW(BD_Obs, BD_Coord, shp, grid_resolution=1, p=2, Mask=False, n_round=1, stat_validation=held)
    train = sorted(n for n in names if n not in held)            # IDW feeds stations Cod-sorted
    ti = [names.index(n) for n in train]
    want = ref.idw_reference(cx, cy, sx[ti], sy[ti], obs[:, ti], 2.0)
    want_r = np.array([[ref.r_round(v, 1) for v in row] for row in want])
    got = out["Ensamble"].values
    # rounding to 1 digit can flip on values within 1e-13 of a tie: allow no more than a handful
    diff = ~(np.isclose(got, want_r, rtol=0, atol=1e-12) | (np.isnan(got) & np.isnan(want_r)))
    assert diff.sum() <= 2
    grid = ref.grid_from_extent(200000.0, 200000.0 + ncol * 1000.0, 9000000.0, 9000000.0 + nrow * 1000.0, 1000.0)
    hi = [names.index(n) for n in names if n in held]           # test_data keeps BD_Obs column order
    wv = ref.validate_reference(grid, got, sx[hi], sy[hi], obs[:, hi])
    tab = out["Validation"]
    assert tab["ID"].tolist() == [names[i] for i in hi]
    for j, row in enumerate(wv):
        for k in ref.GOF_COLUMNS:
            a, b = float(tab[k].iloc[j]), row["gof"][k]
            assert (np.isnan(a) and np.isnan(b)) or a == b, (j, k, a, b)


# ---------------------------------------------------------------- plan API / sharding / errors
def test_plan_api_equals_one_shot_and_date_shards_agree():
    import torch

    cx, cy, sx, sy, obs = synthetic.synthetic_case(40, 25, 50, 520, na_frac=0.02, plant_zero_dist=2, seed=9)
    ref_out = engine.idw_grid(cx, cy, sx, sy, obs)
    plan = engine.Plan(0, cx, cy, sx, sy, 520)
    plan.set_obs(obs)
    out = torch.full((520, cx.size), -1.0, dtype=torch.float64, device="cuda")
    plan.idw(out.data_ptr(), 2.0)
    plan.sync()
    assert np.array_equal(out.cpu().numpy().T, ref_out, equal_nan=True)
    # arbitrary, tile-unaligned date windows give identical values
    out2 = torch.full((520, cx.size), -1.0, dtype=torch.float64, device="cuda")
    for lo, hi in [(0, 7), (7, 301), (301, 520)]:
        plan.idw(out2[lo:].data_ptr(), 2.0, t0=lo, t1=hi)
    plan.sync()
    assert np.array_equal(out2.cpu().numpy().T, ref_out, equal_nan=True)
    assert plan.zero_pairs == 2 and plan.launch_count > 0
    plan.close()
    if torch.cuda.device_count() >= 2:
        multi = engine.idw_grid(cx, cy, sx, sy, obs, devices=[0, 1])
        assert np.array_equal(multi, ref_out, equal_nan=True)


def test_weight_cache_chunking_and_reset(monkeypatch):
    """The A-operand cache is chunked over cell tiles when it does not fit its memory budget; the
    result must not depend on the chunking, nor on whether the cache is reused or regenerated."""
    import torch

    cx, cy, sx, sy, obs = synthetic.synthetic_case(70, 41, 130, 300, na_frac=0.01, plant_zero_dist=3, seed=13)
    want = engine.idw_grid(cx, cy, sx, sy, obs, p=2.0)
    assert_parity(want, ref.idw_reference(cx, cy, sx, sy, obs, 2.0), label="unchunked")
    monkeypatch.setenv("IPR_WCACHE_MB", "1")          # 1 MiB -> a few cell tiles per chunk
    got = engine.idw_grid(cx, cy, sx, sy, obs, p=2.0)
    assert np.array_equal(got, want, equal_nan=True)
    radii = np.array([5e3, 20e3])
    c1 = engine.cressman_grid(cx, cy, sx, sy, obs, radii)
    monkeypatch.delenv("IPR_WCACHE_MB")
    c2 = engine.cressman_grid(cx, cy, sx, sy, obs, radii)
    assert np.array_equal(c1, c2, equal_nan=True)
    # cached vs regenerated weights, and switching weight functions on one plan
    plan = engine.Plan(0, cx, cy, sx, sy, 300)
    plan.set_obs(obs)
    o = torch.empty((3, 300, cx.size), dtype=torch.float64, device="cuda")
    plan.idw(o[0].data_ptr(), 2.0)
    plan.idw(o[1].data_ptr(), 3.0)
    plan.reset_cache()
    plan.idw(o[2].data_ptr(), 2.0)
    plan.sync()
    assert np.array_equal(o[0].cpu().numpy(), o[2].cpu().numpy(), equal_nan=True)
    assert np.array_equal(o[0].cpu().numpy().T, want, equal_nan=True)
    assert_parity(o[1].cpu().numpy().T, ref.idw_reference(cx, cy, sx, sy, obs, 3.0), label="p=3 after p=2")
    plan.close()


@pytest.mark.parametrize("int8", ["1", "0"])
def test_weight_cache_chunking_with_masked_granules(monkeypatch, int8):
    """Heavy NA (masked granules) while the A-operand cache holds only a few cell tiles at a time:
    every chunk of cell tiles must get its masked granules, on the int8 mask path and on the FP64 one."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(70, 41, 130, 300, na_frac=0.25, plant_zero_dist=3, seed=17)
    monkeypatch.setenv("IPR_MASK_INT8", int8)
    want = engine.idw_grid(cx, cy, sx, sy, obs, p=2.0)
    assert_parity(want, ref.idw_reference(cx, cy, sx, sy, obs, 2.0), label="unchunked")
    monkeypatch.setenv("IPR_WCACHE_MB", "1")
    got = engine.idw_grid(cx, cy, sx, sy, obs, p=2.0)
    assert np.array_equal(got, want, equal_nan=True)


@pytest.mark.parametrize("n_dates", [300, 70, 20])
def test_multi_device_scheduler_on_one_gpu(n_dates):
    """The multi-GPU scheduler of the host-buffer entry points (contiguous cell ranges per device,
    2-D copies into the caller's matrix), exercised on a single device by listing it several times
    (one host thread + plan per entry): the shards must tile the matrix exactly like a single-device
    run, for IDW and for every Cressman stack, into pageable (staged) memory."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(25, 20, 45, n_dates, na_frac=0.04, plant_zero_dist=1,
                                                   all_na_dates=(3,), seed=n_dates)
    one = engine.idw_grid(cx, cy, sx, sy, obs, devices=[0])
    for devs in ([0, 0], [0, 0, 0, 0, 0]):
        many = engine.idw_grid(cx, cy, sx, sy, obs, devices=devs)
        # cell tiles start elsewhere in each shard, which changes nothing per cell except the order in
        # which a thread block sums: identical to rounding
        assert_parity(many, one, tol=1e-12, label=f"devices={devs}")
    radii = np.array([4e3, 9e3, 30e3])
    c1 = np.ascontiguousarray(engine.cressman_grid(cx, cy, sx, sy, obs, radii, devices=[0]))
    c3 = np.ascontiguousarray(engine.cressman_grid(cx, cy, sx, sy, obs, radii, devices=[0, 0, 0], n_round=None))
    # a shard's cell tiles are composed differently, so a cell's station list is padded and grouped into
    # k-steps differently: the same products, summed in another order
    assert np.array_equal(np.isnan(c1), np.isnan(c3))
    assert_parity(c3, c1, tol=1e-12, label="sharded cressman")
    assert_parity(one, ref.idw_reference(cx, cy, sx, sy, obs, 2.0), label="sharded idw")


def test_concurrent_one_shot_calls_from_two_threads():
    """include/interpolater_b200.h: the one-shot entry points may be called concurrently (the pools
    behind them are process-wide); each thread gets the result of its own inputs."""
    import threading

    cases = [synthetic.synthetic_case(60, 50, 70, 200, na_frac=0.1, seed=31),
             synthetic.synthetic_case(45, 70, 90, 150, na_frac=0.3, plant_zero_dist=2, seed=32)]
    want = [engine.idw_grid(*c) for c in cases] + [engine.cressman_grid(*cases[0], np.array([8e3, 30e3]))]
    got = [None] * 3
    errors = []

    def work(i):
        try:
            for _ in range(3):
                if i < 2:
                    got[i] = engine.idw_grid(*cases[i])
                else:
                    got[i] = engine.cressman_grid(*cases[0], np.array([8e3, 30e3]))
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    th = [threading.Thread(target=work, args=(i,)) for i in range(3)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    for g, w in zip(got, want):
        assert np.array_equal(g, w, equal_nan=True)


def test_error_paths_report_messages():
    lib = _native.load()
    cx, cy, sx, sy, obs = synthetic.synthetic_case(4, 4, 5, 3)
    with pytest.raises(_native.NativeError, match="p must be a positive"):
        engine.idw_grid(cx, cy, sx, sy, obs, p=-1.0)
    with pytest.raises(_native.NativeError, match="not available"):
        engine.idw_grid(cx, cy, sx, sy, obs, devices=[99])
    bad = obs.copy()
    bad[0, 0] = np.inf
    with pytest.raises(_native.NativeError, match="infinite"):
        engine.idw_grid(cx, cy, sx, sy, bad)
    err = _native.errbuf()
    rc = lib.ipr_idw_f64(None, None, 0, None, None, 0, None, 0, 2.0, None, 0, None, err, _native.ERRLEN)
    assert rc == -1 and b"NULL" in err.value


# ---------------------------------------------------------------- full-size properties (cfg 3)
def test_full_size_properties():
    """1,000,000 cells x 2,000 stations: (i) a field that is constant over stations is reproduced
    at every cell, (ii) linearity in the observations, (iii) min <= out <= max, (iv) a sample of
    cells equals the oracle — checked on the device so that no 29 GB stack crosses PCIe."""
    import torch

    ncol = nrow = 1000
    n_st, n_dates = 2000, 256
    cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, n_st, n_dates)
    const = np.linspace(0.5, 40.0, n_dates)
    z2 = np.repeat(const[:, None], n_st, axis=1)
    plan = engine.Plan(0, cx, cy, sx, sy, n_dates)
    out1 = torch.empty((n_dates, cx.size), dtype=torch.float64, device="cuda")
    out2 = torch.empty_like(out1)
    out3 = torch.empty_like(out1)
    plan.set_obs(obs)
    plan.idw(out1.data_ptr())
    plan.set_obs(z2)
    plan.idw(out2.data_ptr())
    plan.set_obs(2.0 * obs + 3.0 * z2)
    plan.idw(out3.data_ptr())
    plan.sync()
    c = torch.from_numpy(const).cuda()[:, None]
    assert float(((out2 - c).abs() / c).max()) < 1e-12                       # (i)
    lin = 2.0 * out1 + 3.0 * out2
    assert float(((out3 - lin).abs() / lin.abs().clamp_min(1e-12)).max()) < 1e-9   # (ii)
    lo = torch.from_numpy(obs.min(axis=1)).cuda()[:, None]
    hi = torch.from_numpy(obs.max(axis=1)).cuda()[:, None]
    assert bool(((out1 >= lo - 1e-9) & (out1 <= hi + 1e-9)).all())            # (iii)
    idx = np.random.default_rng(0).choice(cx.size, 1500, replace=False)
    want = ref.idw_reference(cx[idx], cy[idx], sx, sy, obs[:6], 2.0)
    got = out1[:6][:, torch.from_numpy(idx).cuda()].cpu().numpy().T
    assert_parity(got, want, label="full-size sample")                      # (iv)
    plan.close()


def test_full_size_properties_masked_and_cressman():
    """The same 1,000,000 x 2,000 geometry with 20 % NA (masked contraction, int8 mask GEMM) and
    through Cressman with the cfg-4 radii: station-constant fields, bounds, NA pattern, linearity,
    and a sample of cells against the oracle — all reduced on the device."""
    import torch

    ncol = nrow = 1000
    n_st, n_dates = 2000, 256
    cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, n_st, n_dates, na_frac=0.2, plant_zero_dist=5,
                                                   all_na_dates=(7,), all_zero_dates=(130,))
    navail = np.isfinite(obs)
    const = np.linspace(0.5, 40.0, n_dates)
    z2 = np.where(navail, const[:, None], np.nan)               # constant over the available stations
    plan = engine.Plan(0, cx, cy, sx, sy, n_dates)
    out1 = torch.empty((n_dates, cx.size), dtype=torch.float64, device="cuda")
    out2 = torch.empty_like(out1)
    plan.set_obs(obs)
    plan.idw(out1.data_ptr())
    plan.set_obs(z2)
    plan.idw(out2.data_ptr())
    plan.sync()
    # --- masked IDW
    assert bool(out1[7].isnan().all()) and bool(out2[7].isnan().all())          # all-NA date
    assert bool((out1[130] == 0).all())                                         # exact-zero layer
    live = [t for t in range(n_dates) if t != 7]
    c = torch.from_numpy(const).cuda()[:, None]
    assert not bool(out2[live].isnan().any())
    assert float(((out2[live] - c[live]).abs() / c[live]).max()) < 1e-9         # constant field
    lo = torch.from_numpy(np.nanmin(obs[live], axis=1)).cuda()[:, None]
    hi = torch.from_numpy(np.nanmax(obs[live], axis=1)).cuda()[:, None]
    assert bool(((out1[live] >= lo - 1e-9) & (out1[live] <= hi + 1e-9)).all())
    idx = np.random.default_rng(1).choice(cx.size, 1200, replace=False)
    dates = [0, 7, 63, 64, 130, 255]
    want = ref.idw_reference(cx[idx], cy[idx], sx, sy, obs[dates], 2.0)
    got = out1[dates][:, torch.from_numpy(idx).cuda()].cpu().numpy().T
    assert_parity(got, want, label="full-size masked sample")
    del out2
    # --- Cressman, cfg-4 radii; the denominator ignores availability, so out <= max and the NA
    # pattern is the same on every date
    radii = np.array([25e3, 50e3, 100e3, 200e3])
    outc = torch.empty((radii.size, n_dates, cx.size), dtype=torch.float64, device="cuda")
    plan.set_obs(obs)
    plan.cressman(outc.data_ptr(), radii)
    plan.sync()
    nanmap = outc.isnan()
    assert bool((nanmap == nanmap[:, :1]).all())
    frac_na = nanmap[:, 0].double().mean(dim=1).cpu().numpy()
    assert frac_na[0] > 0.0 and np.all(np.diff(frac_na) <= 0) and frac_na[-1] == 0.0   # small radii leave holes
    hi_all = torch.from_numpy(np.nanmax(np.where(navail, obs, 0.0), axis=1)).cuda()[None, :, None]
    ok = nanmap | ((outc >= -1e-9) & (outc <= hi_all + 1e-9))
    assert bool(ok.all())
    assert bool((outc[:, 7][~nanmap[:, 7]] == 0).all())                         # all-NA date -> 0 where in range
    wantc = ref.cressman_reference(cx[idx], cy[idx], sx, sy, obs[dates], radii)
    gotc = outc[:, dates][:, :, torch.from_numpy(idx).cuda()].cpu().numpy()
    for ri in range(radii.size):
        assert_parity(gotc[ri].T, wantc[ri], label=f"full-size Cressman sample R={radii[ri]:.0f}")
    plan.close()


# ---------------------------------------------------------------- full records, drop-in signature at full size
def test_full_size_full_record_all_dates_sampled_cells():
    """cfg 3 at its full size AND full length: 1,000,000 cells x 2,000 stations x 3,650 days through the
    29-tile launch path with the ragged 66-date last granule; 1,000 seeded cells x every date against the
    oracle (matrix form, held to the literal per-date loop by tests/test_oracle.py)."""
    import torch

    cx, cy, sx, sy, obs = synthetic.synthetic_case(1000, 1000, 2000, 3650, all_na_dates=(3649,), all_zero_dates=(3600,))
    plan = engine.Plan(0, cx, cy, sx, sy, 3650)
    out = torch.empty((3650, cx.size), dtype=torch.float64, device="cuda")
    plan.set_obs(obs)
    plan.idw(out.data_ptr())
    plan.sync()
    idx = np.sort(np.random.default_rng(5).choice(cx.size, 1000, replace=False))
    got = out[:, torch.from_numpy(idx).cuda()].cpu().numpy().T
    assert_parity(got, ref.idw_reference_matrix(cx[idx], cy[idx], sx, sy, obs, 2.0), label="cfg 3, all 3650 dates")
    assert bool(out[3649].isnan().all()) and bool((out[3600] == 0).all())
    assert bool(torch.isfinite(out[3584:3649]).all())          # the ragged last granule, every cell
    plan.close()


def test_full_size_cfg5_through_the_drop_in_signature():
    """BASELINE.json configs[4] at full size through the reference's own signature: 1,000,000-cell grid,
    2,000 stations with 20 % per-date NA, 400 of them held out as stat_validation, 3,650 days.  The
    validation table must equal what the oracle derives at the 400 test-station cells (IDW at those cells
    from the 1,600 training stations -> R's round -> gof), and a seeded sample of the returned stack must
    match the oracle on every date."""
    import pandas as pd

    ncol = nrow = 1000
    n_st, n_dates = 2000, 3650
    cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, n_st, n_dates, na_frac=0.2, seed=505)
    names = [f"S{i:04d}" for i in range(n_st)]
    dates = pd.date_range("2001-01-01", periods=n_dates, freq="D")
    BD_Obs = pd.concat([pd.DataFrame({"Date": dates}), pd.DataFrame(obs, columns=names)], axis=1)
    BD_Coord = pd.DataFrame({"Cod": names, "X": sx, "Y": sy, "Z": 0.0})
    held_idx = np.sort(np.random.default_rng(20260105).choice(n_st, 400, replace=False))
    held = [names[i] for i in held_idx]
    shp = ipr.SpatVector.from_bbox(200000.0, 200000.0 + ncol * 1000.0, 9000000.0, 9000000.0 + nrow * 1000.0)
    out = ipr.IDW(BD_Obs, BD_Coord, shp, grid_resolution=1, p=2, Mask=False, n_round=1, stat_validation=held)
    ens, tab = out["Ensamble"], out["Validation"]
    assert ens.nlyr() == n_dates and ens.values.shape == (ncol * nrow, n_dates)
    ti = np.array([i for i in range(n_st) if names[i] not in set(held)])       # names are already Cod-sorted
    # (1) a seeded sample of the stack, every date
    idx = np.sort(np.random.default_rng(6).choice(cx.size, 300, replace=False))
    want = ref.idw_reference_matrix(cx[idx], cy[idx], sx[ti], sy[ti], obs[:, ti], 2.0)
    got = ens.values[idx, :]
    want_r = np.round(want, 1)
    close = np.isclose(got, want_r, rtol=0, atol=1e-12) | (np.isnan(got) & np.isnan(want_r))
    # 1-digit rounding may flip on values within 1e-12 of a tie (x.x5): a handful at most, never more than 0.1
    assert (~close).sum() <= 5 and np.nanmax(np.abs(got - want)) <= 0.05 + 1e-9
    # (2) the validation table against the oracle evaluated at the test-station cells
    grid = ref.grid_from_extent(200000.0, 200000.0 + ncol * 1000.0, 9000000.0, 9000000.0 + nrow * 1000.0, 1000.0)
    assert tab["ID"].tolist() == held
    wv = ref.validate_reference(grid, ens.values, sx[held_idx], sy[held_idx], obs[:, held_idx])
    for j, row in enumerate(wv):
        for k in ref.GOF_COLUMNS:
            a, b = float(tab[k].iloc[j]), row["gof"][k]
            assert (np.isnan(a) and np.isnan(b)) or a == b, (j, k, a, b)
    # ... and independently of the engine's stack: oracle IDW at those cells
    cells = np.array(ref.extract_cells(grid, sx[held_idx], sy[held_idx]))
    sim = ref.idw_reference_matrix(cx[cells], cy[cells], sx[ti], sy[ti], obs[:, ti], 2.0)
    eng = ens.values[cells, :]
    assert np.nanmax(np.abs(eng - sim)) <= 0.05 + 1e-9 and np.array_equal(np.isnan(eng), np.isnan(sim))


# ---------------------------------------------------------------- more than one physical GPU
def test_all_visible_gpus_equal_one_device():
    """The product entry point over every visible GPU (one process, a host thread + plan per device, each
    device writing its column range of the caller's matrix) against the one-device run: same cells, same
    arithmetic -> equal to summation order."""
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one visible GPU: the scheduler itself is covered by test_multi_device_scheduler_on_one_gpu")
    cx, cy, sx, sy, obs = synthetic.synthetic_case(300, 257, 333, 700, na_frac=0.08, plant_zero_dist=4,
                                                   all_na_dates=(11,), all_zero_dates=(12,), seed=88)
    devs = list(range(n))
    one = engine.idw_grid(cx, cy, sx, sy, obs, devices=[0])
    many = engine.idw_grid(cx, cy, sx, sy, obs, devices=devs)
    assert np.array_equal(np.isnan(one), np.isnan(many))
    assert_parity(many, one, tol=1e-12, label=f"IDW on devices {devs}")
    b = engine.cell_split_bounds(cx.size, n)
    sample = np.concatenate([np.arange(b[g], min(b[g] + 50, b[g + 1])) for g in range(n)])
    assert_parity(many[sample], ref.idw_reference(cx[sample], cy[sample], sx, sy, obs, 2.0), label="multi-GPU vs oracle")
    radii = np.array([6e3, 25e3, 90e3])
    c1 = np.ascontiguousarray(engine.cressman_grid(cx, cy, sx, sy, obs, radii, devices=[0]))
    cn = np.ascontiguousarray(engine.cressman_grid(cx, cy, sx, sy, obs, radii, devices=devs))
    # a shard composes its cell tiles (hence each station list's padding and k-step grouping) differently
    assert_parity(cn, c1, tol=1e-12, label=f"Cressman on devices {devs}")
    # pinned caller buffer: cudaMemcpy2DAsync straight into the column ranges
    lib = _native.load()
    import ctypes as Ct
    nbytes = cx.size * obs.shape[0] * 8
    hp = lib.ipr_host_alloc(nbytes)
    try:
        buf = np.ctypeslib.as_array((Ct.c_double * (nbytes // 8)).from_address(hp)).reshape(obs.shape[0], cx.size).T
        engine.idw_grid(cx, cy, sx, sy, obs, devices=devs, out=buf)
        assert np.array_equal(buf, many, equal_nan=True)
    finally:
        lib.ipr_host_free(hp)


# ---------------------------------------------------------------- asynchronous jobs (progress, cancel)
def test_job_api_progress_and_result():
    cx, cy, sx, sy, obs = synthetic.synthetic_case(120, 90, 150, 1100, na_frac=0.1, plant_zero_dist=2, seed=21)
    want = engine.idw_grid(cx, cy, sx, sy, obs)
    job = engine.Job.idw(cx, cy, sx, sy, obs)
    seen = []
    while True:
        fin, done, total = job.wait(5)
        seen.append(done)
        if fin:
            break
    assert total == 10 and seen[-1] == total and seen == sorted(seen)   # 9 granules, the first one split in two
    assert np.array_equal(job.finish(), want, equal_nan=True)
    radii = np.array([5e3, 3e4])
    jobc = engine.Job.cressman(cx, cy, sx, sy, obs, radii, devices=[0, 0])
    fin, done, total = jobc.wait(-1)
    assert fin and done == total == 2 * (2 * 9 + 1)
    # two shards compose their cell tiles differently from one: same products, another order of summation
    assert_parity(np.ascontiguousarray(jobc.finish()),
                  np.ascontiguousarray(engine.cressman_grid(cx, cy, sx, sy, obs, radii)), tol=1e-12, label="cressman job")


def test_job_cancel_returns_cancelled_and_leaves_the_library_usable():
    cx, cy, sx, sy, obs = synthetic.synthetic_case(400, 400, 500, 3000, seed=22)
    job = engine.Job.idw(cx, cy, sx, sy, obs)
    job.wait(20)
    job.cancel()
    with pytest.raises(_native.NativeError) as ei:
        job.finish()
    assert ei.value.code == _native.IPR_ECANCELLED
    small = synthetic.synthetic_case(20, 20, 30, 40, seed=23)
    assert_parity(engine.idw_grid(*small), ref.idw_reference(*small, 2.0), label="after a cancelled job")


def test_more_than_65535_colocated_cells():
    """A grid whose stations are snapped to cell centres: more zero-distance cells than one grid.y
    dimension holds.  Every such cell takes its station's observation (running mean of one, R/IDW.R:310-326)."""
    ncol = nrow = 300
    n_st = 70000
    cx, cy = synthetic.regular_grid(ncol, nrow)
    rng = np.random.default_rng(9)
    cells = rng.choice(ncol * nrow, n_st, replace=False)
    obs = np.round(rng.gamma(0.8, 8.0, (3, n_st)), 1) + 0.1
    obs[1, ::7] = np.nan
    got = engine.idw_grid(cx, cy, cx[cells], cy[cells], obs)
    for t in range(3):
        ok = ~np.isnan(obs[t])
        assert np.array_equal(got[cells[ok], t], obs[t, ok])
    assert not np.isnan(got).any()


# ---------------------------------------------------------------- Kriging_Ordinary's IDW fallback
@pytest.mark.parametrize("shape", [(40, 30, 60, 300), (9, 7, 11, 12), (5, 3, 2, 4), (33, 20, 140, 700)])
def test_kriging_idw_fallback_matches_oracle(shape):
    """ipr_idw_kriging_f64 against R/Kriging_Ordinary.R:406-423, :465-476 as restated by the oracle: NA-free and
    NA-bearing granules (FP64 masked GEMM), co-located stations (weight 1e20, present and absent), dates with
    none / one / only equal available values."""
    ncol, nrow, n_st, n_dates = shape
    cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, n_st, n_dates, na_frac=0.2,
                                                   plant_zero_dist=min(3, n_st), seed=n_st)
    obs[0, :] = np.nan
    obs[1, :] = np.nan
    obs[1, n_st // 2] = 4.75
    obs[2, :] = np.where(np.isnan(obs[2, :]), np.nan, 1.5)
    if n_dates > 140:
        obs[140:, :] = np.where(np.isnan(obs[140:, :]), 0.3, obs[140:, :])   # NA-free granules: unmasked tiles
    got = engine.idw_kriging_fallback_grid(cx, cy, sx, sy, obs)
    want = ref.kriging_idw_fallback_reference(cx, cy, sx, sy, obs)
    assert not np.isnan(got).any()
    assert_parity(got, want, label=f"kriging fallback {shape}")
    assert np.all(got[:, 0] == 0.0) and np.all(got[:, 1] == 4.75)
    if np.unique(obs[2][~np.isnan(obs[2])]).size == 1:
        assert np.all(got[:, 2] == 1.5)


def test_kriging_idw_fallback_differs_from_idw_only_where_the_rules_differ():
    """Same contraction as IDW(): away from co-located stations and special dates the two agree to rounding."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(30, 20, 40, 200, na_frac=0.1, seed=5)
    obs = np.where(np.abs(obs) < 1e-12, 0.7, obs)            # no all-zero dates, no constant dates
    a = engine.idw_grid(cx, cy, sx, sy, obs, p=2.0)
    b = engine.idw_kriging_fallback_grid(cx, cy, sx, sy, obs)
    assert_parity(b, a, tol=1e-12, label="kriging fallback vs IDW away from the special cases")


def test_kriging_idw_fallback_multi_shard_and_after_idw():
    """The fallback through the cell-split scheduler, on a plan cache that held IDW weights just before."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(25, 20, 45, 150, na_frac=0.05, plant_zero_dist=2, seed=8)
    engine.idw_grid(cx, cy, sx, sy, obs)
    one = engine.idw_kriging_fallback_grid(cx, cy, sx, sy, obs, devices=[0])
    many = engine.idw_kriging_fallback_grid(cx, cy, sx, sy, obs, devices=[0, 0, 0])
    assert_parity(many, one, tol=1e-12, label="kriging fallback shards")
    assert_parity(one, ref.kriging_idw_fallback_reference(cx, cy, sx, sy, obs), label="kriging fallback")


# ---------------------------------------------------------------- ordinary kriging (the kriging branch)
def _kriging_params(n_dates, rng, models=("exponential", "spherical", "gaussian", "linear", "idw")):
    par = []
    for t in range(n_dates):
        m = models[t % len(models)]
        par.append(dict(use_idw=True) if m == "idw" else
                   dict(nugget=float(rng.uniform(0.5, 2.0)) * (4.0 if m == "gaussian" else 1.0),   # gaussian: ill-conditioned
                        sill=float(rng.uniform(10.0, 60.0)),
                        range=float(rng.uniform(8e3, 40e3)) * (0.3 if m == "gaussian" else 1.0), model=m))
    return par


def _kriging_engine_args(par):
    model = [0 if p.get("use_idw") else ref.VARIOGRAM_MODELS[p["model"]] for p in par]
    get = lambda k: [0.0 if p.get("use_idw") else p[k] for p in par]   # noqa: E731
    rng_ = [1.0 if p.get("use_idw") else p["range"] for p in par]
    return model, get("nugget"), get("sill"), rng_


@pytest.mark.parametrize("shape", [(23, 17, 12, 20), (9, 7, 30, 140), (40, 30, 70, 10), (5, 3, 2, 6)])
def test_kriging_matches_oracle(shape):
    """ipr_kriging_f64 (one LU per date + the cell x station variogram pass) against the reference's literal
    per-cell solve(gamma_matrix, gamma_vector) as restated by the oracle, for the four variogram models, fallback
    dates, NA, constant / single-station / empty dates.  Tolerance: 1e-9, widened by the conditioning of the
    date's gamma matrix (the two solve orders agree to kappa * eps; kappa is asserted to stay moderate)."""
    ncol, nrow, n_st, n_dates = shape
    rng = np.random.default_rng(n_st)
    cx, cy, sx, sy, obs = synthetic.synthetic_case(ncol, nrow, n_st, n_dates, na_frac=0.15, seed=n_st + 1)
    obs = obs + np.where(np.isnan(obs), 0.0, rng.uniform(0.0, 0.5, obs.shape))      # few ties: no constant dates by accident
    obs[0, :] = np.nan
    if n_dates > 3:
        obs[1, :] = np.nan
        obs[1, 0] = 2.5
        obs[2, :] = np.where(np.isnan(obs[2, :]), np.nan, 1.25)
    par = _kriging_params(n_dates, rng)
    model, nug, sill, rng_ = _kriging_engine_args(par)
    got = engine.kriging_grid(cx, cy, sx, sy, obs, model, nug, sill, rng_)
    want = ref.kriging_reference(cx, cy, sx, sy, obs, par)
    assert not np.isnan(got).any() and np.all(got >= 0.0)
    assert np.all(got[:, 0] == 0.0)
    if n_dates > 3:
        assert np.all(got[:, 1] == 2.5) and np.all(got[:, 2] == 1.25)
    dss = ref.distance_matrix(sx, sy, sx, sy)
    n_ill = 0
    for t in range(n_dates):
        tol = 1e-9
        p = par[t]
        ok = ~np.isnan(obs[t])
        if not p.get("use_idw") and ok.sum() >= 2:
            n = int(ok.sum())
            gm = np.full((n + 1, n + 1), p["nugget"])
            off = ~np.eye(n, dtype=bool)
            gm[:n, :n][off] = ref.variogram_models(dss[np.ix_(ok, ok)], p["nugget"], p["sill"], p["range"], p["model"])[off]
            gm[n, :n] = gm[:n, n] = 1.0
            gm[n, n] = 0.0
            kappa = np.linalg.cond(gm)
            if kappa > 1e9:      # near the reference's own kappa > 1e12 fallback: the two solve orders are not comparable
                n_ill += 1
                continue
            tol = max(1e-9, 100 * kappa * 2.0 ** -53)
        # absolute floor on the scale of the observations: max(0, .) clips values that cancel to ~0
        scale = np.nanmax(np.abs(obs[t])) if ok.any() else 1.0
        assert np.all(np.abs(got[:, t] - want[:, t]) <= tol * np.maximum(np.abs(want[:, t]), scale)), (t, par[t])
    assert n_ill <= n_dates // 4, "most dates must be well enough conditioned to be compared"


def test_kriging_singular_system_takes_the_mean_and_shards_agree():
    """Two stations at the same place make gamma_matrix exactly singular: solve() raises in the reference and the
    layer is mean(available_values) (R/Kriging_Ordinary.R:494).  Also: the cell-split scheduler on the kriging path."""
    cx, cy, sx, sy, obs = synthetic.synthetic_case(12, 9, 8, 4, seed=3)
    obs = obs + 0.1 * np.arange(obs.shape[1])[None, :] + 0.3
    sx[5], sy[5] = sx[2], sy[2]                                        # duplicate location -> two equal rows
    par = [dict(nugget=1.0, sill=20.0, range=15e3, model="exponential")] * 4
    model, nug, sill, rng_ = _kriging_engine_args(par)
    one = engine.kriging_grid(cx, cy, sx, sy, obs, model, nug, sill, rng_, devices=[0])
    for t in range(4):
        assert np.allclose(one[:, t], obs[t].mean(), rtol=1e-13, atol=0)
    cx, cy, sx, sy, obs = synthetic.synthetic_case(30, 21, 25, 150, na_frac=0.1, seed=4)
    rng = np.random.default_rng(1)
    par = _kriging_params(150, rng)
    model, nug, sill, rng_ = _kriging_engine_args(par)
    one = engine.kriging_grid(cx, cy, sx, sy, obs, model, nug, sill, rng_, devices=[0])
    many = engine.kriging_grid(cx, cy, sx, sy, obs, model, nug, sill, rng_, devices=[0, 0, 0])
    assert np.array_equal(one, many)          # the solve does not depend on the cell range; the pass is per cell
