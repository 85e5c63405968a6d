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
