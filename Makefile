# Convenience targets; __graft_entry__.build() is the canonical build (same nvcc command).
PYTHON ?= python
NVCC   ?= /usr/local/cuda/bin/nvcc
LIB     = interpolater_b200/lib/libinterpolater_b200.so
SRC     = interpolater_b200/csrc/ipr_api.cu
HDR     = interpolater_b200/csrc/ipr_kernels.cuh include/interpolater_b200.h

.PHONY: lib test test-gpu bench clean

lib: $(LIB)

$(LIB): $(SRC) $(HDR)
	mkdir -p $(dir $(LIB))
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC \
	        -I include -o $@ $(SRC)

test: lib
	$(PYTHON) -m pytest tests -q -m "not gpu"

test-gpu: lib
	$(PYTHON) -m pytest tests -q -m gpu

bench: lib
	$(PYTHON) bench.py

clean:
	rm -f $(LIB)
