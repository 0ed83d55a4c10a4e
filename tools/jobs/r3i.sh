#!/bin/bash
O=gpurun_out/r3i; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
tl() {
  B200SP_LIB_SUFFIX=_tl timeout 120 python tools/tcg2_timeline.py $1 $2 $3 > $O/tl_$2_$(echo $1 | tr , x).txt 2>&1
}
tl 602112,16,32 wgrad 30
tl 150528,24,144 wgrad 30
tl 602112,96,16 wgrad 30
tl 150528,24,144 dgrad 20
head -40 $O/tl_wgrad_602112x16x32.txt
