#!/bin/bash
O=gpurun_out/r4a; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_presplit_gpu.py tests/test_krn_gpu.py tests/test_dann_gpu.py tests/test_krn_tf32_gpu.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline 2>> $O/bench.err | head -c 200 | grep -o '"ms_per_step": [0-9.]*'; done
