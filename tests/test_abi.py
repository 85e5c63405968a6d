"""The C-ABI library loads on a CPU-only box and exports every symbol the header declares
(no compute calls here)."""
import os
import re

from conftest import ROOT
from interpolater_b200 import _native


def _declared_functions():
    with open(os.path.join(ROOT, "include", "interpolater_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ipr_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared_functions() == sorted(_native.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = _native.load()
    for name in _declared_functions():
        assert hasattr(lib, name), f"{name} is declared in include/interpolater_b200.h but not exported"
    assert lib.ipr_version() >= 100


def test_na_real_bit_pattern():
    import struct

    lib = _native.load()
    assert struct.pack("<d", lib.ipr_na_real()) == struct.pack("<Q", 0x7FF00000000007A2)


def test_product_code_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under interpolater_b200/ may reference it."""
    pkg = os.path.join(ROOT, "interpolater_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                with open(os.path.join(dirpath, fn), errors="replace") as f:
                    txt = f.read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{fn} imports oracle"


def test_cell_split_bounds_of_the_library_match_the_python_twin():
    """ipr_cell_split_bounds is what run_host_job splits a multi-GPU job with; sharding.cell_block_bounds is the
    twin bench.py's strong-scaling arm uses.  Contiguous, monotone, on 64-cell tiles, equal to within a tile."""
    import random

    from interpolater_b200 import engine, sharding

    rng = random.Random(7)
    cases = [(1, 1), (63, 2), (64, 2), (65, 2), (1_000_000, 8), (1_000_000, 7), (100, 16)]
    cases += [(rng.randint(1, 3_000_000), rng.randint(1, 16)) for _ in range(300)]
    for n, g in cases:
        b = engine.cell_split_bounds(n, g)
        assert b == sharding.cell_block_bounds(n, g)
        assert b[0] == 0 and b[-1] == n and all(x <= y for x, y in zip(b, b[1:]))
        assert all(x % 64 == 0 for x in b[:-1])
        sizes = [y - x for x, y in zip(b, b[1:])]
        assert max(sizes) - min(sizes) <= 64 + 63      # one tile, plus the ragged last tile


def test_async_job_fails_cleanly_without_a_device():
    """The job API on a CPU-only box: start succeeds (the work happens on a worker thread), wait reports
    completion, finish returns the CUDA error — no crash, no hang, nothing computed on the CPU."""
    import pytest

    from interpolater_b200 import engine, synthetic

    lib = _native.load()
    if lib.ipr_device_count() >= 1:
        pytest.skip("a device is present: covered by the GPU tests")
    cx, cy, sx, sy, obs = synthetic.synthetic_case(6, 5, 7, 9)
    job = engine.Job.idw(cx, cy, sx, sy, obs)
    fin, done, total = job.wait(-1)
    assert fin and done == 0 and total == 1
    with pytest.raises(_native.NativeError):
        job.finish()
