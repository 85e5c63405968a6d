#!/bin/bash
# one 1-GPU batch: tests, Cressman A/B of the A-prefetch depth, the default bench line, sanitizer check
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q > $O/r02_pytest_gpu_e.log 2>&1; tail -3 $O/r02_pytest_gpu_e.log
for v in 1 2; do
  echo "== APREF=$v"; IPR_LIB_OVERRIDE=$PWD/interpolater_b200/lib/libipr_apref$v.so timeout 200 python tools/perf_cressman.py 2>&1 | sed -n 2,3p
done > $O/r02_cr_apref.log 2>&1; cat $O/r02_cr_apref.log
timeout 600 python bench.py > $O/r02_bench_default_c.json 2> $O/r02_bench_default_c.err; tail -c 300 $O/r02_bench_default_c.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_default_c.json").read().strip().splitlines()[-1])
print("idw", d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["ms_per_step"], d.get("parity_sample_ok"))
for k, v in d["workloads"].items():
    print(k, v["ms_per_step"], v["roofline"]["frac"], v["e2e"]["ms_per_step"], v.get("parity_sample_ok"))
PY
timeout 400 bash tools/asan_check.sh
