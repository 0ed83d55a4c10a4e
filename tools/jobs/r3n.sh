#!/bin/bash
O=gpurun_out/r3n; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
run() { echo "== $*"; env "$@" timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline 2>> $O/bench.err | head -c 200 | grep -o '"ms_per_step": [0-9.]*'; }
run B200SP_TCG2_LEAN=off
run B200SP_TCG2_LEAN=dgrad B200SP_TCG2_LEAN_R=96
run B200SP_TCG2_LEAN=all B200SP_TCG2_LEAN_R=96
run B200SP_TCG2_LEAN=all B200SP_TCG2_LEAN_R=192
run B200SP_TCG2_LEAN=all B200SP_TCG2_LEAN_R=576
run B200SP_TCG2_LEAN=all B200SP_TCG2_LEAN_R=100000
run B200SP_TCG2_LEAN=all B200SP_TCG2_LEAN_R=96 B200SP_TCG2_EW=8
