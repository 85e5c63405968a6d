import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA sm_100 device (run with -m gpu on a B200)")


def pytest_collection_modifyitems(config, items):
    """Plain `pytest` on a box without an sm_100 device skips the gpu-marked tests instead of erroring on
    every one of them.  When the GPU tests were asked for explicitly (`-m gpu`) nothing is skipped: a missing
    device or library must fail loudly there — the engine itself has no fallback."""
    markexpr = config.getoption("-m", default="") or ""
    if "gpu" in markexpr and "not gpu" not in markexpr:
        return
    try:
        from interpolater_b200 import _native

        have = _native.load().ipr_device_count() >= 1
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no sm_100 CUDA device (the engine has no CPU fallback); run with -m gpu on a B200")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def assert_parity(got, want, tol=1e-9, label=""):
    """The parity bar of SURVEY §8c: NA positions identical; |got - ref| <= tol * |ref| elementwise
    (absolute floor tol * 1e-12 for exact zeros)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{label}: shape {got.shape} != {want.shape}"
    na_g, na_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(na_g, na_w), f"{label}: NA pattern differs ({na_g.sum()} vs {na_w.sum()} NA)"
    m = ~na_w
    err = np.abs(got[m] - want[m])
    lim = tol * np.maximum(np.abs(want[m]), 1e-12)
    bad = err > lim
    assert not bad.any(), (f"{label}: {bad.sum()} values off; max rel err "
                           f"{np.max(err / np.maximum(np.abs(want[m]), 1e-300)):.3e} > {tol}")


@pytest.fixture(scope="session")
def packaged():
    """The reference's packaged fixture (BD_Obs / BD_Coord / study_area.shp) as frozen in
    tests/golden/packaged_inputs.npz, rebuilt as the objects the public API takes."""
    import pandas as pd

    from interpolater_b200.spatial import SpatVector

    z = np.load(os.path.join(GOLDEN, "packaged_inputs.npz"))
    stations = [str(s) for s in z["stations"]]
    dates = pd.to_datetime(z["dates"], unit="D", origin="unix")
    BD_Obs = pd.DataFrame({"Date": dates, **{s: z["obs"][:, i] for i, s in enumerate(stations)}})
    BD_Coord = pd.DataFrame({"Cod": [str(c) for c in z["cod"]], "X": z["X"], "Y": z["Y"], "Z": z["Z"]})
    bbox = tuple(float(v) for v in z["bbox"])
    shp = SpatVector([z["ring"]], crs=str(z["prj"]), bbox=bbox)
    return dict(BD_Obs=BD_Obs, BD_Coord=BD_Coord, shapefile=shp, raw=z, stations=stations)


@pytest.fixture(scope="session")
def packaged_golden():
    return np.load(os.path.join(GOLDEN, "packaged_golden.npz"))


@pytest.fixture(scope="session")
def synthetic_golden():
    return np.load(os.path.join(GOLDEN, "synthetic_golden.npz"))


RAIN_THRESHOLD = dict(no_rain=(0, 1), light_rain=(1, 5), moderate_rain=(5, 20), heavy_rain=(20, 40),
                      extremely_rain=(40, float("inf")))
