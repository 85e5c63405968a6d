#!/bin/bash
O=gpurun_out
timeout 200 python -m pytest tests -m gpu -q -k "all_visible or multi_device or date_shards" > $O/r02_pytest_8gpu_b.log 2>&1; tail -2 $O/r02_pytest_8gpu_b.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > $O/r02_bench_n8_strong_b.json 2> $O/r02_bench_n8_strong_b.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n8_strong_b.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "inproc", (d.get("e2e_inproc") or {}).get("ms_per_step"), d.get("parity_sample_ok"))
for k, v in d["workloads"].items():
    print(k, v["value"], v["ms_per_step"], "e2e ms", v["e2e"]["ms_per_step"], v.get("parity_sample_ok"))
PY
