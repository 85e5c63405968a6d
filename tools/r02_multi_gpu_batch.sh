#!/bin/bash
# multi-GPU batch (run under gpurun --gpus N): topology, host-sink probe, multi-GPU tests, strong-scaling bench,
# one-process device-list probe
N=${1:-8}
O=gpurun_out
nvidia-smi topo -m > $O/r02_topo_${N}gpu.txt 2>&1
timeout 300 python tools/d2h_probe.py > $O/r02_d2h_probe_${N}gpu.log 2>&1; cp $O/d2h_probe.json $O/r02_d2h_probe_${N}gpu.json
timeout 200 python -m pytest tests -m gpu -q -k "all_visible or multi_device or date_shards" > $O/r02_pytest_${N}gpu.log 2>&1; tail -2 $O/r02_pytest_${N}gpu.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 > $O/r02_bench_n${N}_strong.json 2> $O/r02_bench_n${N}_strong.err
tail -c 400 $O/r02_bench_n${N}_strong.err
timeout 300 python tools/inproc_probe.py $N > $O/r02_inproc_probe_${N}gpu.log 2> $O/r02_inproc_probe_${N}gpu.err
IPR_D2H_ROWS=1 timeout 300 python tools/inproc_probe.py $N > $O/r02_inproc_probe_rows_${N}gpu.log 2> $O/r02_inproc_probe_rows_${N}gpu.err
grep -h "GB/s\|call" $O/r02_inproc_probe_${N}gpu.log $O/r02_inproc_probe_rows_${N}gpu.log
python - $N <<'PY'
import json, sys
n = sys.argv[1]
d = json.loads(open(f"gpurun_out/r02_bench_n{n}_strong.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e ms", d["e2e"]["ms_per_step"], "inproc", (d.get("e2e_inproc") or {}).get("ms_per_step"))
for k, v in d["workloads"].items():
    print(k, v["value"], v["ms_per_step"], "e2e ms", v["e2e"]["ms_per_step"])
print(json.load(open(f"gpurun_out/r02_d2h_probe_{n}gpu.json"))["d2h_gbs"])
PY
