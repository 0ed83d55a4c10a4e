#!/bin/bash
O=gpurun_out/r3h; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
S="602112,16,32 602112,96,16 150528,24,96 150528,144,24 150528,24,144 37632,32,144 37632,192,32 37632,32,192 9408,576,96 9408,96,576 9408,384,64 9408,64,384"
for g in 1 2 4; do
B200SP_TCG2_WG_GROUPS=$g timeout 300 python tools/gemm_bench.py --graph --ops wgrad --shapes $S > $O/wg$g.txt 2>&1
done
paste $O/wg1.txt $O/wg2.txt $O/wg4.txt | awk '{print $1,$2,$3,$8,$13}'
B200SP_TCG2_WG_GROUPS=2 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "wgrad" 2>&1 | tail -2
