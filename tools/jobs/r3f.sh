#!/bin/bash
O=gpurun_out/r3f; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "pw_" > $O/tests.txt 2>&1; echo "tests rc=$?" >> $O/tests.txt
tail -3 $O/tests.txt
S="602112,16,32 602112,96,16 150528,24,96 150528,144,24 150528,24,144 37632,32,144 37632,192,32 37632,32,192 9408,576,96 9408,96,576 9408,384,64 9408,64,384"
for ew in 8 0; do
B200SP_TCG2_EW=$ew B200SP_TCG2_EW_R=100000 timeout 300 python tools/gemm_bench.py --graph --ops fwd,dgrad --shapes $S > $O/gemm_ew$ew.txt 2>&1
done
paste $O/gemm_ew8.txt $O/gemm_ew0.txt | awk '{print $1,$2,$3,$8}'
for r in 96 192 100000; do
echo "EW_R=$r"; B200SP_TCG2_EW_R=$r timeout 600 python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline 2> $O/bench.err | head -c 330
done > $O/bench_sweep.txt; cat $O/bench_sweep.txt
