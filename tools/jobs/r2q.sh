#!/bin/bash
O=gpurun_out/r2q; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
for s in "9408,96,576 fwd 12" "2352,1024,1280 fwd 12" "9408,576,96 wgrad 12"; do
  set -- $s
  B200SP_LIB_SUFFIX=_tl timeout 120 python tools/tcg2_timeline.py $1 $2 $3 > $O/tl_$2_$(echo $1 | tr , x).txt 2>&1
done
