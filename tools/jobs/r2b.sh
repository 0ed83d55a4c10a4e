#!/bin/bash
# round-2 GPU job B: first run of the second-generation tcgen05 GEMM (tcgemm2.cu, env B200SP_TCG2=1)
O=gpurun_out/r2b; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
B200SP_TCG2=1 timeout 600 python -m pytest -q tests/test_kernels_gpu.py -k "pw_fwd or pw_dgrad or pw_wgrad" 2>&1 | tail -60 > $O/t1_tcg2_kernels.log
B200SP_TCG2=1 B200SP_TCG2_WGRAD=0 timeout 600 python -m pytest -q tests/test_kernels_gpu.py -k "pw_fwd or pw_dgrad or pw_wgrad" 2>&1 | tail -30 > $O/t2_tcg2_nowgrad_kernels.log
B200SP_TCG2=1 timeout 600 python -m pytest -q -x tests/test_krn_gpu.py tests/test_dann_gpu.py 2>&1 | tail -15 > $O/t3_tcg2_models.log
B200SP_TCG2=1 timeout 600 python tools/gemm_bench.py --reps 5 > $O/gemm_bench_tcg2.txt 2>&1
timeout 600 python tools/gemm_bench.py --reps 5 > $O/gemm_bench_r1.txt 2>&1
B200SP_TCG2=1 timeout 600 python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof_tcg2.txt > $O/bench_tcg2.json 2> $O/bench_tcg2.err
B200SP_TCG2=1 B200SP_TCG2_WGRAD=0 timeout 600 python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof_tcg2_nowgrad.txt > $O/bench_tcg2_nowgrad.json 2> $O/bench_tcg2_nowgrad.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof_r1.txt > $O/bench_r1.json 2> $O/bench_r1.err
ls -la $O
