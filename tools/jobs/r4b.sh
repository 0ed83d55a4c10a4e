#!/bin/bash
O=gpurun_out/r4b; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline 2>> $O/bench.err | head -c 200 | grep -o '"ms_per_step": [0-9.]*'; }
run B200SP_X=1
run B200SP_TCG2_PRE_M=10000
run B200SP_WGDIRECT_MIN_M=30000
