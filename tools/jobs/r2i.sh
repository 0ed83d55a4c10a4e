#!/bin/bash
# round-2 GPU job I: programmatic dependent launch on the GEMM, merged TMEM waits / L2 prefetch in the epilogue, torch ops
O=gpurun_out/r2i; mkdir -p $O
export B200SP_NO_AUTOBUILD=1 B200SP_TCG2=1
timeout 600 python -m pytest -q tests/test_kernels_gpu.py tests/test_torch_ops_gpu.py tests/test_krn_gpu.py 2>&1 | tail -30 > $O/t_tests.log
timeout 600 python tools/gemm_bench.py --reps 5 > $O/gemm_bench.txt 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof_pdl1.txt > $O/bench_pdl1.json 2> $O/bench_pdl1.err
B200SP_PDL=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof_pdl0.txt > $O/bench_pdl0.json 2> $O/bench_pdl0.err
