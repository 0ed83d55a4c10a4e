#!/bin/bash
# final-state ncu --set full captures: long-M data gradient (576-thread split) and the stride-1 depthwise backward (prefetch)
O=gpurun_out/r4c; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
N="timeout 150 ncu --set full --clock-control none --import-source on"
$N -k regex:tcgemm2 -s 3 -c 1 -o $O/dgrad_150528x24x144_final python tools/gemm_bench.py --reps 2 --ops dgrad --shapes 150528,24,144 > $O/a.log 2>&1
$N -k regex:dwr_bwd2 -s 24 -c 1 -o $O/dwbwd_final python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline > $O/b.log 2>&1
ls -la $O
