#!/bin/bash
O=gpurun_out/r2l; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
S="--shapes 2352,1024,1280 9408,576,96 602112,96,16 150528,144,24 2352,320,960"
timeout 300 python tools/gemm_bench.py --reps 7 --ops wgrad $S > $O/wgrad_head.txt 2>&1
B200SP_PDL=0 timeout 300 python tools/gemm_bench.py --reps 7 --ops wgrad $S > $O/wgrad_head_pdl0.txt 2>&1
B200SP_LIB_SUFFIX=_nm timeout 300 python tools/gemm_bench.py --reps 7 --ops wgrad $S > $O/wgrad_nomerge.txt 2>&1
B200SP_LIB_SUFFIX=_e B200SP_TCG2=1 timeout 300 python tools/gemm_bench.py --reps 7 --ops wgrad $S > $O/wgrad_r2e.txt 2>&1
