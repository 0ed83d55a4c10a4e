#!/bin/bash
O=gpurun_out/r3v; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_krn_gpu.py tests/test_dann_gpu.py -m gpu -x -q -k "dw_ or train or step or dann" 2>&1 | tail -4
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline 2>> $O/bench.err | head -c 200 | grep -o '"ms_per_step": [0-9.]*'; }
run B200SP_DW_SLAB=0
run B200SP_DW_SLAB=1
