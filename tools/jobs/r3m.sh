#!/bin/bash
O=gpurun_out/r3m; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
S="602112,16,32 602112,96,16 150528,24,96 150528,144,24 150528,24,144 37632,32,144 37632,192,32 37632,32,192 9408,64,384 9408,96,576"
timeout 300 python tools/gemm_bench.py --graph --ops fwd,dgrad --shapes $S > $O/base.txt 2>&1
B200SP_TCG2_LEAN=all timeout 300 python tools/gemm_bench.py --graph --ops fwd,dgrad --shapes $S > $O/lean.txt 2>&1
paste $O/base.txt $O/lean.txt | awk '{print $1,$2,$3,$8}'
B200SP_TCG2_LEAN=all timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "pw_fwd or pw_dgrad" 2>&1 | tail -2
