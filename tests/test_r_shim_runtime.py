"""The .Call shim (r-package/src/r_shim.c), linked against the real C-ABI library and EXECUTED under
a minimal mock of R's runtime (tests/mock_r/mock_runtime.c plays R: builds the SEXPs, looks the
routine up in the registration table, calls it, catches `error()`).

CPU: without a device the core fails and the shim must turn that into an R error, with the protect
stack balanced.  GPU: the arrays that come back through the shim match the oracle."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_parity
from interpolater_b200 import _native, synthetic

pytestmark = pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")


@pytest.fixture(scope="module")
def driver(tmp_path_factory):
    libdir = os.path.dirname(os.path.abspath(_native.LIB_PATH))
    exe = str(tmp_path_factory.mktemp("shim") / "shim_driver")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Wno-cast-function-type",
           "-I", os.path.join(ROOT, "tests", "mock_r"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "mock_r", "mock_runtime.c"), os.path.join(ROOT, "r-package", "src", "r_shim.c"),
           "-L", libdir, "-linterpolater_b200", "-Wl,-rpath," + libdir, "-o", exe]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    return exe


def _write_case(path, mode, cx, cy, sx, sy, obs, last):
    with open(path, "wb") as f:
        f.write(struct.pack("<5q", mode, cx.size, sx.size, obs.shape[0], len(last)))
        for a in (cx, cy, sx, sy, np.asfortranarray(obs).ravel(order="F"), np.asarray(last, np.float64)):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())


def _run(driver, tmp_path, mode, case, last, **env):
    cx, cy, sx, sy, obs = case
    src, dst = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    _write_case(src, mode, cx, cy, sx, sy, obs, last)
    if os.path.exists(dst):
        os.remove(dst)
    e = dict(os.environ)
    e.update({k: str(v) for k, v in env.items()})
    res = subprocess.run([driver, src, dst], capture_output=True, text=True, timeout=300, env=e)
    out = np.fromfile(dst, dtype=np.float64) if res.returncode == 0 else None
    return res, out


def _has_gpu():
    return _native.load().ipr_device_count() > 0


def test_shim_raises_an_r_error_without_a_device(driver, tmp_path):
    if _has_gpu():
        pytest.skip("a device is present: covered by the GPU test")
    case = synthetic.synthetic_case(6, 5, 7, 9, na_frac=0.1)
    for mode, last in ((0, [2.0]), (1, [5e3, 2e4]), (2, []), (3, [1.0] * 9 + [1.0] * 9 + [20.0] * 9 + [9e3] * 9)):
        res, out = _run(driver, tmp_path, mode, case, last, MOCK_PROGRESS=1)
        assert res.returncode == 3, (res.returncode, res.stderr)      # error() was raised, nothing crashed
        assert out is None and "R error: interpolater_b200 (code -" in res.stderr


@pytest.mark.gpu
def test_shim_end_to_end_matches_oracle(driver, tmp_path):
    from oracle import reference_port as ref

    case = synthetic.synthetic_case(37, 23, 41, 150, na_frac=0.15, plant_zero_dist=2, all_na_dates=(4,),
                                    all_zero_dates=(9,), seed=77)
    cx, cy, sx, sy, obs = case
    want_idw = ref.idw_reference(cx, cy, sx, sy, obs, 2.0)
    for pinned in (0, 1):     # R's own (pageable) allocation, then the cudaHostAlloc-backed R_allocator_t
        res, out = _run(driver, tmp_path, 0, case, [2.0], MOCK_PINNED=pinned, MOCK_PROGRESS=1)
        assert res.returncode == 0, res.stderr
        got = out.reshape(obs.shape[0], cx.size).T                    # column-major n_cells x n_dates
        assert_parity(got, want_idw, label=f"IDW through the .Call shim (pinned={pinned})")
        na_bits = np.frombuffer(got[np.isnan(got)].tobytes(), dtype=np.uint64)
        assert na_bits.size and np.all(na_bits == 0x7FF00000000007A2)  # is.na() and printing as NA in R
        assert f"pinned allocs {pinned} frees {pinned}" in res.stderr  # the GC hands the pages back through mem_free
        # progress(done, total) ran on the main thread, monotonically, and reached the total (3 work items)
        seen = [tuple(map(float, l.split(":")[1].split("/"))) for l in res.stderr.splitlines() if l.startswith("progress ")]
        assert seen and seen[-1][0] == seen[-1][1] == 3.0 and seen == sorted(seen)
    radii = [4e3, 1.5e4]
    res, out = _run(driver, tmp_path, 1, case, radii)
    assert res.returncode == 0, res.stderr
    stacks = out.reshape(len(radii), obs.shape[0], cx.size)
    want = ref.cressman_reference(cx, cy, sx, sy, obs, np.asarray(radii))
    for ri in range(len(radii)):
        assert_parity(stacks[ri].T, want[ri], label=f"Cressman R={radii[ri]:.0f} through the .Call shim")
    res, out = _run(driver, tmp_path, 2, case, [])
    assert res.returncode == 0, res.stderr
    assert_parity(out.reshape(obs.shape[0], cx.size).T, ref.kriging_idw_fallback_reference(cx, cy, sx, sy, obs),
                  label="Kriging IDW fallback through the .Call shim")
    # ordinary kriging: model codes + fitted parameters per date (a third of the dates fall back to IDW)
    kc = synthetic.synthetic_case(19, 13, 14, 30, na_frac=0.1, seed=78)
    D = kc[4].shape[0]
    par = [dict(use_idw=True) if t % 3 == 0 else dict(nugget=1.0 + 0.01 * t, sill=25.0, range=12e3, model=("exponential", "spherical")[t % 2])
           for t in range(D)]
    codes = [0.0 if p.get("use_idw") else float(ref.VARIOGRAM_MODELS[p["model"]]) for p in par]
    extra = codes + [p.get("nugget", 0.0) for p in par] + [p.get("sill", 0.0) for p in par] + [p.get("range", 1.0) for p in par]
    res, out = _run(driver, tmp_path, 3, kc, extra)
    assert res.returncode == 0, res.stderr
    want = ref.kriging_reference(kc[0], kc[1], kc[2], kc[3], kc[4], par)
    got = out.reshape(D, kc[0].size).T
    assert np.all(np.abs(got - want) <= 1e-8 * np.maximum(np.abs(want), np.nanmax(kc[4])))


@pytest.mark.gpu
def test_shim_interrupt_and_failing_progress_cancel_the_job(driver, tmp_path):
    """A user interrupt (R_CheckUserInterrupt inside R_ToplevelExec) or an error in the progress callback cancels
    the GPU job; the R error is raised only after the workers were joined; the protect stack stays balanced."""
    case = synthetic.synthetic_case(300, 300, 300, 3000, seed=5)       # long enough to be caught in flight
    res, out = _run(driver, tmp_path, 0, case, [2.0], MOCK_INTERRUPT_AT=2, MOCK_PROGRESS=1)
    assert res.returncode == 3 and "interrupted by the user" in res.stderr, res.stderr
    res, out = _run(driver, tmp_path, 0, case, [2.0], MOCK_PROGRESS=1, MOCK_PROGRESS_FAIL_AT=1)
    assert res.returncode == 3 and "progress callback failed" in res.stderr, res.stderr
    # and the library is usable afterwards (same process would be; here: a fresh run completes)
    small = synthetic.synthetic_case(9, 8, 7, 20, seed=6)
    res, out = _run(driver, tmp_path, 0, small, [2.0])
    assert res.returncode == 0, res.stderr
