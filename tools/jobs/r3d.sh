#!/bin/bash
O=gpurun_out/r3d; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
N="timeout 200 ncu --set full --clock-control none --import-source on -k regex:tcgemm2 -s 3 -c 1"
$N -o $O/dgrad_150528x24x144 python tools/gemm_bench.py --reps 2 --ops dgrad --shapes 150528,24,144 > $O/a.log 2>&1
$N -o $O/fwd_150528x144x24 python tools/gemm_bench.py --reps 2 --ops fwd --shapes 150528,144,24 > $O/b.log 2>&1
ls -la $O
