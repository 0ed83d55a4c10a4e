#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
export B200SP_NO_AUTOBUILD=1 B200SP_TCG2=1
for s in "602112,96,16 fwd 12" "150528,24,144 dgrad 12" "602112,16,32 fwd 12" "9408,96,576 fwd 4"; do
  set -- $s
  B200SP_LIB_SUFFIX=_tl timeout 120 python tools/tcg2_timeline.py $1 $2 $3 > $O/tl_$2_$(echo $1 | tr , x).txt 2>&1
done
timeout 600 python -m pytest -q tests/test_krn_gpu.py -k bs48 -s 2>&1 | tail -60 > $O/t_bs48.log
