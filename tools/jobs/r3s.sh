#!/bin/bash
O=gpurun_out/r3s; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_presplit_gpu.py -m gpu -x -q -k "dgrad" 2>&1 | tail -2
S="602112,16,32 602112,96,16 150528,24,96 150528,144,24 150528,24,144 37632,32,144 37632,32,192 9408,96,576 2352,1024,1280"
timeout 300 python tools/gemm_bench.py --graph --ops dgrad --shapes $S 2>&1 | tail -9
timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline 2>> $O/bench.err | head -c 200 | grep -o '"ms_per_step": [0-9.]*'
