#!/bin/bash
# round-2 GPU job T: A operand through tensor memory (TS form of tcgen05.mma), env B200SP_TCG2_TS=1
O=gpurun_out/r2t; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
B200SP_TCG2_TS=1 timeout 300 python -m pytest -q tests/test_kernels_gpu.py -k "pw_fwd or pw_dgrad" 2>&1 | tail -40 > $O/t_kernels_ts.log
B200SP_TCG2_TS=1 timeout 300 python tools/gemm_bench.py --reps 5 --ops fwd,dgrad > $O/gemm_bench_ts.txt 2>&1
B200SP_TCG2_TS=1 timeout 300 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof_ts.txt > $O/bench_ts.json 2> $O/bench_ts.err
B200SP_TCG2_TS=1 B200SP_LIB_SUFFIX=_tl timeout 120 python tools/tcg2_timeline.py 9408,96,576 fwd 12 > $O/tl_fwd_9408x96x576.txt 2>&1
