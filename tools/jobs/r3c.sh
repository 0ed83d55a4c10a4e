#!/bin/bash
O=gpurun_out/r3c; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "dgrad" > $O/tests.txt 2>&1; echo "tests rc=$?" >> $O/tests.txt
tail -3 $O/tests.txt
S="602112,16,32 602112,96,16 150528,24,96 150528,144,24 150528,24,144 37632,32,144 37632,32,192 9408,96,576"
timeout 300 python tools/gemm_bench.py --graph --ops dgrad --shapes $S > $O/gemm_dgrad.txt 2>&1; cat $O/gemm_dgrad.txt
