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
    assert set(registered) == {"C_ipr_idw", "C_ipr_cressman"}
    helper = open(os.path.join(ROOT, "r-package", "R", "ipr_core.R")).read()
    for name, nargs in registered.items():
        m = re.search(r"\.Call\(%s,(.*?)\n\s*if \(is\.null\(devices\)\)" % name, helper, flags=re.S)
        assert m, f"{name} is not called from ipr_core.R"
        # arguments before the trailing `devices` expression
        args = [a for a in re.split(r",(?![^()]*\))", m.group(1)) if a.strip()]
        assert len(args) + 1 == int(nargs), (name, args)
    # balanced delimiters (cheap guard against a truncated edit; R itself is not available)
    for o, c in ("()", "{}", "[]"):
        assert helper.count(o) == helper.count(c)
