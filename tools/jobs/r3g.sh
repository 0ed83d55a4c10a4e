#!/bin/bash
O=gpurun_out/r3g; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_presplit_gpu.py -m gpu -x -q -k "pw_ or fwd or dgrad or wgrad" > $O/tests.txt 2>&1; echo "tests rc=$?" >> $O/tests.txt
tail -3 $O/tests.txt
S="602112,16,32 602112,96,16 150528,24,96 150528,144,24 150528,24,144 37632,32,144 37632,192,32 37632,32,192 9408,576,96 9408,96,576 2352,960,160 2352,1024,1280"
timeout 300 python tools/gemm_bench.py --graph --shapes $S > $O/gemm.txt 2>&1; cat $O/gemm.txt
for i in 1 2; do timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline 2> $O/bench.err | head -c 330; echo; done > $O/bench2.txt; cat $O/bench2.txt
