#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
B200SP_LIB_SUFFIX=_tl timeout 120 python tools/tcg2_timeline.py 9408,96,576 fwd 8 > $O/tl_imm.txt 2>&1
B200SP_LIB_SUFFIX=_nr timeout 120 python tools/tcg2_timeline.py 9408,96,576 fwd 8 > $O/tl_imm_norecheck.txt 2>&1
timeout 300 python tools/gemm_bench.py --reps 5 > $O/gemm_bench.txt 2>&1
timeout 300 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench.json 2> $O/bench.err
