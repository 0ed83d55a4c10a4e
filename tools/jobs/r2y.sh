#!/bin/bash
O=gpurun_out/r2y; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 900 python -m pytest tests/test_presplit_gpu.py -m gpu -x -q > $O/tests.txt 2>&1; echo "tests rc=$?" >> $O/tests.txt
tail -5 $O/tests.txt
S="2352,160,576 2352,960,160 2352,160,960 2352,320,960 2352,1024,320 2352,1024,1024 2352,1024,1280"
timeout 300 python tools/gemm_bench.py --shapes $S > $O/gemm_pre.txt 2>&1
B200SP_WS=0 timeout 300 python tools/gemm_bench.py --shapes $S > $O/gemm_gen.txt 2>&1
paste $O/gemm_pre.txt $O/gemm_gen.txt | cut -c1-120
timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench.json 2> $O/bench.err; tail -c 1500 $O/bench.json
