#!/bin/bash
O=gpurun_out/r3j; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "wgrad" > $O/tests.txt 2>&1; echo "tests rc=$?" >> $O/tests.txt
tail -5 $O/tests.txt
S="602112,16,32 602112,96,16 150528,24,96 150528,144,24 150528,24,144 37632,32,144 37632,192,32 37632,32,192"
timeout 300 python tools/gemm_bench.py --graph --ops wgrad --shapes $S > $O/wg_direct.txt 2>&1
B200SP_WGDIRECT=0 timeout 300 python tools/gemm_bench.py --graph --ops wgrad --shapes $S > $O/wg_tc.txt 2>&1
paste $O/wg_direct.txt $O/wg_tc.txt | awk '{print $1,$2,$3,$4,$8}'
timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline 2> $O/bench.err | head -c 330
