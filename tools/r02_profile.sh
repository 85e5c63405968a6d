#!/bin/bash
# Round-2 ncu evidence, one GPU (run under gpurun): launch lists of the bench command per workload
# (metrics-only pass), one --set full capture per dominant kernel, DRAM traffic per launch.
# Each ncu pass runs only after the same command has exited 0 without ncu.
set -u
O=gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity"
for wl in idw cressman idw_masked; do
  $B --workload $wl > $O/r02_plain_$wl.json 2> $O/r02_plain_$wl.err || { echo "plain $wl failed"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_$wl.csv \
      $B --workload $wl > $O/r02_ncu_launches_$wl.log 2>&1
done
python tools/ncu_cressman.py 25e3 50e3 100e3 200e3 > $O/r02_plain_ncu_cressman.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cressman_kernel -c 4 -o $O/r02_ncu_cressman -f \
    python tools/ncu_cressman.py 25e3 50e3 100e3 200e3 > $O/r02_ncu_cressman.log 2>&1
python tools/perf_int8_mask.py 1 1 > $O/r02_plain_int8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:denom_umma_kernel -c 1 -o $O/r02_ncu_denom_umma -f \
    python tools/perf_int8_mask.py 1 1 > $O/r02_ncu_denom_umma.log 2>&1
# DRAM bytes of every dominant launch of one masked step (numerator GEMM + mask GEMM)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:"contract_kernel|denom_umma_kernel" -c 8 --csv --log-file $O/r02_traffic_idw_masked.csv \
    python tools/perf_int8_mask.py 1 1 > $O/r02_ncu_traffic_masked.log 2>&1
ls -la $O | grep r02_ | tail -20
