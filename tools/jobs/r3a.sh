#!/bin/bash
O=gpurun_out/r3a; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
S="2352,960,160 2352,320,960 2352,1024,1280"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches.csv python tools/gemm_bench.py --reps 3 --shapes $S > $O/ncu.log 2>&1
for bn in 128 112 96 64; do
  echo "== BN cap $bn"; B200SP_TCG2_PRE_BN=$bn timeout 300 python tools/gemm_bench.py --graph --shapes $S 2>&1 | tail -9
done > $O/bn_sweep.txt 2>&1
cat $O/bn_sweep.txt
