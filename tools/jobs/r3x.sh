#!/bin/bash
O=gpurun_out/r3x; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_krn_gpu.py tests/test_dann_gpu.py tests/test_presplit_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/step_profile.txt 2> $O/bench.err | head -c 330
echo; head -12 $O/step_profile.txt
