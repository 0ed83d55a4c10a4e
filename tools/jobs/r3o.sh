#!/bin/bash
O=gpurun_out/r3o; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline 2>> $O/bench.err | head -c 200 | grep -o '"ms_per_step": [0-9.]*'; }
run B200SP_X=1
run B200SP_TCG2_PW=4
run B200SP_TCG2_LEAN_WG=1
run B200SP_TCG2_LEAN_WG=1 B200SP_TCG2_PW=4
B200SP_TCG2_PW=4 B200SP_TCG2_LEAN_WG=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "pw_" 2>&1 | tail -2
