#!/bin/bash
# round-2 GPU job E: elect.sync MMA/TMA issue (no waterfall loops), elected pollers; full GPU suite with the v2 GEMM on
O=gpurun_out/r2e; mkdir -p $O
export B200SP_NO_AUTOBUILD=1 B200SP_TCG2=1
timeout 900 python -m pytest -q tests -m gpu 2>&1 | tail -25 > $O/t_all_gpu.log
timeout 600 python tools/gemm_bench.py --reps 5 > $O/gemm_bench_tcg2.txt 2>&1
timeout 600 python bench.py --steps 50 --warmup 5 --no-reference-cuda --no-cpu-baseline --profile-out $O/prof_tcg2.txt > $O/bench_tcg2.json 2> $O/bench_tcg2.err
B200SP_TCG2=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof_r1.txt > $O/bench_r1.json 2> $O/bench_r1.err
ls -la $O
