"""CPU tests of the host-side mirror (spatial stand-ins, validation module, argument checks of
IDW()/Cressman()) against the oracle and against the reference's own test expectations
(tests/testthat/test-IDW.R, test-Cressman.R: error strings, layer counts, file creation)."""
import math

import numpy as np
import pandas as pd
import pytest

from conftest import RAIN_THRESHOLD, assert_parity
import interpolater_b200 as ipr
from interpolater_b200 import sharding, spatial, validation
from oracle import reference_port as ref


# ----------------------------------------------------------------------------- spatial
def test_grid_geometry_matches_oracle(packaged):
    shp = packaged["shapefile"]
    g = spatial.GridGeometry.from_extent(shp.ext(), 5000.0, shp.crs)
    o = ref.grid_from_extent(*shp.ext(), 5000.0)
    X, Y = g.cell_centres()
    assert (g.ncol, g.nrow) == (o["ncol"], o["nrow"]) == (4, 6)
    assert np.array_equal(X, o["X"]) and np.array_equal(Y, o["Y"])
    for res in (1000.0, 2500.0, 7000.0, 50000.0):
        g = spatial.GridGeometry.from_extent(shp.ext(), res)
        o = ref.grid_from_extent(*shp.ext(), res)
        assert (g.ncol, g.nrow, g.xmax, g.ymax) == (o["ncol"], o["nrow"], o["xmax"], o["ymax"])


def test_cell_from_xy_matches_oracle(packaged):
    shp = packaged["shapefile"]
    g = spatial.GridGeometry.from_extent(shp.ext(), 5000.0)
    o = ref.grid_from_extent(*shp.ext(), 5000.0)
    px, py = packaged["raw"]["X"], packaged["raw"]["Y"]
    got = g.cell_from_xy(px, py)
    assert got.tolist() == ref.extract_cells(o, px, py)
    assert got[0] == 5 * 4 + 2      # M001 sits in row 5, col 2
    assert got[1] == -1             # M002 lies outside the raster -> extract gives NA


def test_shapefile_reader_roundtrip(tmp_path, packaged):
    """vect() reads a polygon .shp + .prj; written here from the frozen ring of study_area.shp."""
    import struct

    ring = packaged["raw"]["ring"]
    xmin, xmax, ymin, ymax = packaged["shapefile"].ext()
    n = len(ring)
    rec = struct.pack("<i4d2i i", 5, xmin, ymin, xmax, ymax, 1, n, 0) + ring.astype("<f8").tobytes()
    body = struct.pack(">2i", 1, len(rec) // 2) + rec
    hdr = struct.pack(">7i", 9994, 0, 0, 0, 0, 0, (100 + len(body)) // 2) + struct.pack("<2i", 1000, 5)
    hdr += struct.pack("<8d", xmin, ymin, xmax, ymax, 0, 0, 0, 0)
    p = tmp_path / "area.shp"
    p.write_bytes(hdr + body)
    (tmp_path / "area.prj").write_text(str(packaged["raw"]["prj"]))
    v = ipr.vect(str(p))
    assert v.ext() == pytest.approx((xmin, xmax, ymin, ymax))
    assert len(v.rings) == 1 and np.array_equal(v.rings[0], ring)
    assert not v.is_lonlat() and "UTM" in v.crs
    assert spatial.SpatVector([ring], crs='GEOGCS["WGS 84",DATUM["WGS_1984"]]').is_lonlat()


def test_polygon_cover_and_crop_mask():
    g = spatial.GridGeometry(0, 100, 0, 100, 10, 10)
    sq = spatial.SpatVector([np.array([[20, 20], [20, 70], [70, 70], [70, 20], [20, 20]], float)])
    cover = g.polygon_cover(sq).reshape(10, 10)
    assert cover.sum() == 25 and cover[3:8, 2:7].all()   # rows count from the top: y 70..20 -> rows 3..7
    r = spatial.SpatRaster(g, np.asfortranarray(np.arange(200, dtype=float).reshape(100, 2)), ["a", "b"])
    m = r.crop_mask(sq)
    assert (m.geom.ncol, m.geom.nrow) == (5, 5) and m.nlyr() == 2
    assert (m.geom.xmin, m.geom.xmax, m.geom.ymin, m.geom.ymax) == (20, 70, 20, 70)
    assert not np.isnan(m.values).any()
    assert m.layer(0)[0, 0] == r.layer(0)[3, 2]
    tri = spatial.SpatVector([np.array([[0, 0], [100, 0], [0, 100], [0, 0]], float)])
    c = g.polygon_cover(tri).reshape(10, 10)
    assert c[9, 0] and not c[0, 9] and c.sum() == 45


def test_rounding_rules():
    x = np.array([0.125, 0.135, 2.675, -1.005, 1.0, np.nan, 0.0, 12345.678])
    want = np.array([ref.r_round(v, 2) for v in x])
    got = spatial.r_round(x, 2)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])
    assert spatial.terra_round(np.array([0.125, -0.125, 2.5]), 2).tolist() == [0.13, -0.13, 2.5]
    rng = np.random.default_rng(0)
    v = rng.gamma(0.8, 8.0, 2000)
    assert np.array_equal(spatial.r_round(v, 1), np.array([ref.r_round(t, 1) for t in v]))


def test_write_cdf(tmp_path):
    from scipy.io import netcdf_file

    g = spatial.GridGeometry(0, 30, 0, 20, 3, 2, crs="LOCAL")
    vals = np.asfortranarray(np.array([[1, 2], [3, np.nan], [5, 6], [7, 8], [9, 10], [11, 12]], float))
    r = spatial.SpatRaster(g, vals, ["2015-01-01", "2015-01-02"])
    fn = tmp_path / "m.nc"
    r.write_cdf(str(fn))
    with netcdf_file(str(fn), "r", mmap=False) as nc:
        v = nc.variables["precipitation"]
        assert v.shape == (2, 2, 3)
        assert v[0, 0, 0] == 1 and v[1, 1, 2] == 12 and v[1, 0, 1] == pytest.approx(v._FillValue)


# ----------------------------------------------------------------------------- validation
def test_r_sample_matches_r():
    """`set.seed(123); sample(1:10)` is 3 10 2 8 6 9 1 7 5 4 in R >= 3.6 (rejection sampling);
    `set.seed(123); runif(3)` is 0.2875775 0.7883051 0.4089769."""
    assert validation.RMersenneTwister(123).sample_int(10, 10) == [3, 10, 2, 8, 6, 9, 1, 7, 5, 4]
    r = validation.RMersenneTwister(123)
    assert [round(r.unif_rand(), 7) for _ in range(3)] == [0.2875775, 0.7883051, 0.4089769]


def test_select_data_manual_and_random(packaged, capfd):
    obs, crd = packaged["BD_Obs"], packaged["BD_Coord"]
    d = validation.select_data(obs, crd, training=1, stat_validation="M001")
    assert list(d["train_data"].columns) == ["Date"] + [f"M{i:03d}" for i in range(2, 11)]
    assert list(d["test_data"].columns) == ["Date", "M001"]
    assert d["test_cords"]["Cod"].tolist() == ["M001"] and len(d["train_cords"]) == 9
    assert "Manual validation mode. Stations: M001 will be used for validation." in capfd.readouterr().err
    d = validation.select_data(obs, crd, training=0.8)
    # sample(columns, 8) after set.seed(123): indices 3 10 2 8 6 9 1 7
    assert list(d["train_data"].columns)[1:] == ["M003", "M010", "M002", "M008", "M006", "M009", "M001", "M007"]
    assert list(d["test_data"].columns) == ["Date", "M004", "M005"]
    assert "Random validation mode. 80 % of the stations will be used for training and 20 % for validation." in \
        capfd.readouterr().err
    with pytest.raises(ValueError, match="between 0 and 1"):
        validation.select_data(obs, crd, training=1.5)


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_metrics_match_oracle(seed):
    rng = np.random.default_rng(seed)
    n = 80
    obs = np.where(rng.random(n) < 0.5, 0.0, np.round(rng.gamma(0.8, 8.0, n), 1))
    sim = np.abs(obs + rng.normal(0, 2.0, n))
    obs[rng.random(n) < 0.1] = np.nan
    sim[rng.random(n) < 0.1] = np.nan
    got = validation.metrics(sim, obs, RAIN_THRESHOLD)
    want = ref.gof_rounded(sim, obs)
    for k in ref.GOF_COLUMNS:
        a, b = float(got["gof"][k].iloc[0]), want[k]
        assert (math.isnan(a) and math.isnan(b)) or a == b, k
    wc = ref.categorical_metrics(sim, obs, RAIN_THRESHOLD)
    gc = got["categorical_metrics"]
    assert gc["Category"].tolist() == [r["Category"] for r in wc]
    for i, row in enumerate(wc):
        for k in ("POD", "SR", "CSI", "HB", "FAR", "HK", "HSS", "ETS"):
            a, b = float(gc[k].iloc[i]), row[k]
            assert (math.isnan(a) and math.isnan(b)) or a == pytest.approx(b, rel=1e-12), (row["Category"], k)


def test_validate_against_oracle_on_golden_stack(packaged, packaged_golden):
    g = packaged_golden
    shp = packaged["shapefile"]
    geom = spatial.GridGeometry.from_extent(shp.ext(), 5000.0, shp.crs)
    labels = [str(d)[:10] for d in packaged["BD_Obs"]["Date"]]
    stack = spatial.SpatRaster(geom, np.asfortranarray(g["idw"]), labels)
    d = validation.select_data(packaged["BD_Obs"], packaged["BD_Coord"], 1, "M001")
    tab = validation.validate(d["test_cords"], d["test_data"], stack)
    assert tab["ID"].tolist() == ["M001"]
    got = tab[ref.GOF_COLUMNS].to_numpy()[0]
    assert np.array_equal(got, g["idw_gof"][0])
    # with Rain_threshold -> list(gof=, categorical_metrics=)
    stack_c = spatial.SpatRaster(geom, np.asfortranarray(g["cressman"][1]), labels)
    res = validation.validate(d["test_cords"], d["test_data"], stack_c, RAIN_THRESHOLD)
    assert set(res) == {"gof", "categorical_metrics"}
    assert np.array_equal(res["gof"][ref.GOF_COLUMNS].to_numpy()[0], g["cressman_gof"][1, 0])
    cat = res["categorical_metrics"][["POD", "SR", "CSI", "HB", "FAR", "HK", "HSS", "ETS"]].to_numpy()
    assert_parity(cat, g["cressman_cat"][1, 0], tol=1e-12, label="categorical")


# ----------------------------------------------------------------------------- argument checks
# (the reference's error-path tests: tests/testthat/test-IDW.R:44-102, test-Cressman.R:87-169)
def test_idw_error_strings(packaged):
    obs, crd, shp = packaged["BD_Obs"], packaged["BD_Coord"], packaged["shapefile"]
    with pytest.raises(TypeError, match="shapefile must be a 'SpatVector' object."):
        ipr.IDW(obs, crd, "not a vector", grid_resolution=5)
    with pytest.raises(TypeError, match="BD_Obs must be a 'data.table' or a 'data.frame'."):
        ipr.IDW([1, 2], crd, shp, grid_resolution=5)
    with pytest.raises(TypeError, match="BD_Coord must be a 'data.table' or a 'data.frame'."):
        ipr.IDW(obs, [1, 2], shp, grid_resolution=5)
    bad = crd.copy()
    bad.loc[0, "Cod"] = "x"
    with pytest.raises(ValueError, match="The names of the coordinates do not appear in the observed data."):
        ipr.IDW(obs, bad, shp, grid_resolution=5)
    with pytest.raises(TypeError, match="'grid_resolution' must be a single numeric value in km."):
        ipr.IDW(obs, crd, shp, grid_resolution="5")
    ll = spatial.SpatVector(shp.rings, crs='GEOGCS["WGS 84"]')
    with pytest.raises(ValueError, match="projected CRS"):
        ipr.IDW(obs, crd, ll, grid_resolution=5)


def test_cressman_error_strings(packaged):
    obs, crd, shp = packaged["BD_Obs"], packaged["BD_Coord"], packaged["shapefile"]
    with pytest.raises(TypeError, match="shapefile must be a 'SpatVector' with a defined CRS."):
        ipr.Cressman(obs, crd, "nope", grid_resolution=5, search_radius=10)
    with pytest.raises(TypeError, match="BD_Obs must be a 'data.table' or a 'data.frame'."):
        ipr.Cressman([1], crd, shp, grid_resolution=5, search_radius=10)
    with pytest.raises(TypeError, match="BD_Coord must be a 'data.table' or a 'data.frame'."):
        ipr.Cressman(obs, [1], shp, grid_resolution=5, search_radius=10)
    bad = crd.copy()
    bad.loc[0, "Cod"] = "x"
    with pytest.raises(ValueError, match="do not appear.*x"):
        ipr.Cressman(obs, bad, shp, grid_resolution=5, search_radius=10)
    with pytest.raises(TypeError, match="'search_radius' must be a numeric vector"):
        ipr.Cressman(obs, crd, shp, grid_resolution=5, search_radius="a")
    with pytest.raises(TypeError, match="'grid_resolution' must be a single numeric value"):
        ipr.Cressman(obs, crd, shp, grid_resolution=None, search_radius=10)


# ----------------------------------------------------------------------------- sharding
def test_cell_block_bounds():
    for n, g in [(1_000_000, 8), (24, 8), (64, 2), (65, 2), (1, 4), (1000, 3)]:
        b = sharding.cell_block_bounds(n, g)
        assert b[0] == 0 and b[-1] == n and len(b) == g + 1
        assert all(b[i] <= b[i + 1] for i in range(g)) and all(x % 64 == 0 for x in b[:-1])
    b = sharding.cell_block_bounds(1_000_000, 8)
    assert max(b[i + 1] - b[i] for i in range(8)) - min(b[i + 1] - b[i] for i in range(8)) <= 64


def test_date_block_bounds():
    for n, g in [(3650, 8), (3650, 1), (60, 8), (128, 2), (129, 2), (1, 4), (1000, 3)]:
        b = sharding.date_block_bounds(n, g)
        assert b[0] == 0 and b[-1] == n and len(b) == g + 1
        assert all(b[i] <= b[i + 1] for i in range(g))
        assert all(x % 32 == 0 for x in b[:-1])
    b = sharding.date_block_bounds(3650, 8)
    assert b == [0, 448, 896, 1376, 1824, 2272, 2752, 3200, 3650]
    assert max(b[i + 1] - b[i] for i in range(8)) <= 480     # balanced to within one 32-date step


# ----------------------------------------------------------------------------- ingestion
def _write_station_csvs(folder):
    """Four station files shaped like the reference's inst/extdata/Folds_ejs_create_data
    (quoted header, ISO dates, one value column), with ragged date ranges and gaps."""
    days = pd.date_range("2014-12-28", "2015-03-05", freq="D")
    rng = np.random.default_rng(7)
    series = {}
    for k, name in enumerate(["M001", "M002", "M003", "M004"]):
        sel = days[k * 3: len(days) - k * 2]
        vals = np.round(rng.gamma(0.8, 8.0, len(sel)) * (rng.random(len(sel)) < 0.5), 1)
        txt = [f'"Date","{name}"']
        for d, v in zip(sel, vals):
            miss = name == "M003" and d.day % 5 == 0      # 20 % gaps in M003
            txt.append(f"{d.date()},{'NA' if miss else v}")
        (folder / f"{name}.csv").write_text("\n".join(txt) + "\n")
        series[name] = dict(zip([str(d.date()) for d in sel], vals))
    (folder / "notes.txt").write_text("ignored\n")
    return series


def test_create_data_matches_reference_expectations(tmp_path):
    """tests/testthat/test-create_data.R:8-38 — a table without max.na, a (data, Na_stations) pair
    with it, and the same table when the files are read in parallel."""
    from interpolater_b200 import ingest
    series = _write_station_csvs(tmp_path)
    out = ingest.create_data(tmp_path, Start_date="2015-01-01", End_Date="2015-03-01", ncores=None, max_na=None)
    assert isinstance(out, pd.DataFrame)
    assert list(out.columns) == ["Date", "M001", "M002", "M003", "M004"]
    assert len(out) == 60 and out["Date"].is_monotonic_increasing
    assert str(out["Date"].iloc[0])[:10] == "2015-01-01" and str(out["Date"].iloc[-1])[:10] == "2015-03-01"
    for name, rec in series.items():
        for d, v in zip(out["Date"], out[name]):
            key = str(d)[:10]
            if name == "M003" and int(key[-2:]) % 5 == 0:
                assert math.isnan(v)
            elif key in rec:
                assert v == rec[key]
            else:
                assert math.isnan(v)            # outside that station's record -> NA (all = TRUE)
    par = ingest.create_data(tmp_path, Start_date="2015-01-01", End_Date="2015-03-01", ncores=2)
    pd.testing.assert_frame_equal(out, par)

    flt = ingest.create_data(tmp_path, Start_date="2015-01-01", End_Date="2015-03-01", max_na=10)
    assert set(flt) == {"data", "Na_stations"}
    pct = flt["Na_stations"].iloc[0]
    assert pct["M001"] == 0.0 and pct["M003"] == pytest.approx(100.0 * 13 / 60)
    assert "M003" not in flt["data"].columns and "M001" in flt["data"].columns
    # strict "<": a station at exactly the threshold is dropped (R/create_data.R:89)
    edge = ingest.create_data(tmp_path, Start_date="2015-01-01", End_Date="2015-03-01", max_na=100.0 * 13 / 60)
    assert "M003" not in edge["data"].columns

    dates, cols, z = ingest.obs_matrix(out)
    assert z.flags.f_contiguous and z.shape == (60, 4) and cols == ["M001", "M002", "M003", "M004"]


def test_create_data_rejects_wrong_column_count(tmp_path):
    from interpolater_b200 import ingest
    (tmp_path / "M009.csv").write_text("Date,a,b\n2015-01-01,1,2\n")
    with pytest.raises(ValueError, match="2 names"):
        ingest.create_data(tmp_path, "2015-01-01", "2015-01-31")
