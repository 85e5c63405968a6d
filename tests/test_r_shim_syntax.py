"""R is not installed here, so the .Call shim cannot be built; it is at least syntax- and
type-checked against a minimal mock of R's C API and against the real C-ABI header, and the R-side
helper is checked to call exactly the routines the shim registers."""
import os
import re
import shutil
import subprocess

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_r_shim_compiles_against_mock_r_api():
    cmd = ["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-Wno-cast-function-type", "-fsyntax-only",
           "-I", os.path.join(ROOT, "tests", "mock_r"), "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "r-package", "src", "r_shim.c")]
    res = subprocess.run(cmd, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr


def test_r_helper_calls_registered_routines():
    shim = open(os.path.join(ROOT, "r-package", "src", "r_shim.c")).read()
    registered = dict(re.findall(r'\{"(C_ipr_\w+)",\s*\(DL_FUNC\)&\w+,\s*(\d+)\}', shim))
    assert registered == {"C_ipr_idw": "9", "C_ipr_cressman": "9", "C_ipr_idw_kriging": "8", "C_ipr_kriging": "10",
                          "C_ipr_release": "0"}
    helper = open(os.path.join(ROOT, "r-package", "R", "ipr_core.R")).read()
    helper_code = "\n".join(l.split("#")[0] for l in helper.splitlines())
    calls = re.findall(r"\.Call\((C_ipr_\w+)((?:[^()]|\((?:[^()]|\([^()]*\))*\))*)\)", helper_code)
    assert {c[0] for c in calls} == set(registered), "every registered routine is called from ipr_core.R"
    for name, rest in calls:
        # top-level commas of the argument list = number of arguments after the routine symbol
        depth, n = 0, 0
        for ch in rest:
            depth += ch in "([{"
            depth -= ch in ")]}"
            n += (ch == "," and depth == 0)
        assert n == int(registered[name]), (name, n, rest)
    # balanced delimiters (cheap guard against a truncated edit; R itself is not available)
    for o, c in ("()", "{}", "[]"):
        assert helper_code.count(o) == helper_code.count(c)
