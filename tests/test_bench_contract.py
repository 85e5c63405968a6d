"""bench.py's reference arm runs on the host only, so its JSON contract can be checked here."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         env=e, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    return res.stdout.strip().splitlines()


def test_reference_arm_prints_one_json_line():
    lines = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--stations", "60",
                  "--cpu-sample-cells", "400", "--cpu-sample-dates", "4"])
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cell-days/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["metric"].startswith("interpolated cell-days/sec")
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "cell-days/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["scaling"] == "strong"
    # default --workload all: cfg 4 / cfg 5 ride along as sub-records
    assert set(d["workloads"]) == {"cressman", "idw_masked"}
    for sub in d["workloads"].values():
        assert sub["value"] > 0 and sub["cpu_baseline"]["kind"] == "port" and "workload" in sub["config"]


def test_single_workload_has_no_sub_records():
    lines = _run(["--impl", "reference", "--workload", "cressman", "--steps", "1", "--warmup", "0", "--stations", "60",
                  "--cpu-sample-cells", "400", "--cpu-sample-dates", "4"])
    d = json.loads(lines[0])
    assert "workloads" not in d and "Cressman" in d["config"]["workload"] and d["config"]["n_stacks"] == 4


def test_reference_arm_other_ranks_exit_quietly():
    lines = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                 env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert lines == []
