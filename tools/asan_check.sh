#!/bin/bash
# Host-side memory / UB check of the engine (sanitizers).  compute-sanitizer is closed on this pool, so
# the device side rests on the parity suite; this builds the library and the mock-R shim driver with
# -fsanitize=address,undefined and runs IDW, Cressman and the Kriging fallback through the .Call shim on a GPU box
# (ASan and CUDA coexist with protect_shadow_gap=0; leak detection off: the CUDA runtime keeps its own arenas).
#   gpurun -- bash tools/asan_check.sh        -> gpurun_out/r02_asan.log
set -u
T=$(mktemp -d)
O=gpurun_out/r02_asan.log
nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
     -Xcompiler -fsanitize=address -Xcompiler -fsanitize=undefined -Xcompiler -fno-omit-frame-pointer -I include \
     -o $T/libinterpolater_b200.so interpolater_b200/csrc/ipr_api.cu > $O 2>&1 || { echo "asan build failed" >> $O; exit 1; }
gcc -std=c11 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -Wno-cast-function-type \
    -I tests/mock_r -I include tests/mock_r/mock_runtime.c r-package/src/r_shim.c \
    -L $T -linterpolater_b200 -Wl,-rpath,$T -o $T/shim_driver >> $O 2>&1 || { echo "driver build failed" >> $O; exit 1; }
python - "$T" <<'PY' >> $O 2>&1
import struct, sys, numpy as np
sys.path.insert(0, ".")
from interpolater_b200 import synthetic
T = sys.argv[1]
cx, cy, sx, sy, obs = synthetic.synthetic_case(70, 41, 130, 300, na_frac=0.15, plant_zero_dist=3, all_na_dates=(4,), seed=3)
for mode, last in ((0, [2.0]), (1, [4e3, 1.5e4, 6e4]), (2, [])):
    with open(f"{T}/in{mode}.bin", "wb") as f:
        f.write(struct.pack("<5q", mode, cx.size, sx.size, obs.shape[0], len(last)))
        for a in (cx, cy, sx, sy, np.asfortranarray(obs).ravel(order="F"), np.asarray(last, np.float64)):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
print("inputs written")
PY
export ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:abort_on_error=0:halt_on_error=0
export UBSAN_OPTIONS=print_stacktrace=1
rc_all=0
for mode in 0 1 2; do
  for pinned in 0 1; do
    MOCK_PINNED=$pinned MOCK_PROGRESS=1 $T/shim_driver $T/in$mode.bin $T/out$mode.bin >> $O 2>&1
    rc=$?
    echo "mode $mode pinned $pinned -> exit $rc" >> $O
    [ $rc -ne 0 ] && rc_all=1
  done
done
MOCK_INTERRUPT_AT=1 $T/shim_driver $T/in0.bin $T/out0.bin >> $O 2>&1; echo "interrupt -> exit $? (3 expected)" >> $O
n=$(grep -c -E "ERROR: AddressSanitizer|runtime error:" $O)
echo "sanitizer reports: $n ; overall rc $rc_all" >> $O
tail -12 $O
