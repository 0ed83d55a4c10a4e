#!/bin/bash
O=gpurun_out/r3l; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 1500 python -m pytest tests/test_kernels_gpu.py tests/test_presplit_gpu.py tests/test_krn_gpu.py tests/test_krn_tf32_gpu.py tests/test_dann_gpu.py -m gpu -x -q > $O/tests.txt 2>&1; echo "tests rc=$?" >> $O/tests.txt
tail -5 $O/tests.txt
timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline 2> $O/bench.err | head -c 330
