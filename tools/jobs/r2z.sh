#!/bin/bash
O=gpurun_out/r2z; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
S="2352,160,576 2352,960,160 2352,160,960 2352,320,960 2352,1024,320 2352,1024,1024 2352,1024,1280 9408,576,96 9408,96,576 9408,384,64 9408,64,384"
timeout 300 python tools/gemm_bench.py --graph --shapes $S > $O/gemm_pre.txt 2>&1
B200SP_WS=0 timeout 300 python tools/gemm_bench.py --graph --shapes $S > $O/gemm_gen.txt 2>&1
B200SP_TCG2_PRE_M=10000 timeout 300 python tools/gemm_bench.py --graph --shapes $S > $O/gemm_pre10k.txt 2>&1
paste $O/gemm_pre.txt $O/gemm_gen.txt $O/gemm_pre10k.txt | awk '{print $1,$2,$3,$8,$13}'
