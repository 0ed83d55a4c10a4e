#!/bin/bash
# round-2 GPU job J: v2 GEMM as default, single-pass TF32 (--use_fp16) mode, torch ops, full GPU suite
O=gpurun_out/r2j; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 1200 python -m pytest -q tests -m gpu -s 2>&1 | tail -60 > $O/t_all_gpu.log
timeout 600 python tools/gemm_bench.py --reps 5 > $O/gemm_bench.txt 2>&1
timeout 900 python bench.py --steps 50 --warmup 5 --no-reference-cuda --no-cpu-baseline --profile-out $O/prof.txt > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --steps 50 --warmup 5 --dtype tf32 --no-secondary --no-cpu-baseline --profile-out $O/prof_tf32.txt > $O/bench_tf32.json 2> $O/bench_tf32.err
B200SP_TCG2_WGRAD=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_r1wgrad.json 2> $O/bench_r1wgrad.err
