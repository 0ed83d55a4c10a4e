#!/bin/bash
O=gpurun_out/r2o; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 600 python -m pytest -q tests/test_kernels_gpu.py -k "pw_" 2>&1 | tail -5 > $O/t_kernels.log
timeout 600 python tools/gemm_bench.py --reps 5 > $O/gemm_bench.txt 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof.txt > $O/bench.json 2> $O/bench.err
B200SP_TCG2_WGRAD=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_r1wgrad.json 2> $O/bench_r1wgrad.err
