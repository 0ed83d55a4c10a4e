#!/bin/bash
# round-2 GPU job F: clock64 timelines of the v2 GEMM (debug library) + the tests that failed in r2e + new parity tests
O=gpurun_out/r2f; mkdir -p $O
export B200SP_NO_AUTOBUILD=1 B200SP_TCG2=1
for s in "2352,1024,1280 fwd 30" "9408,96,576 fwd 24" "602112,96,16 fwd 40" "150528,24,144 dgrad 40" "602112,16,32 wgrad 30" "9408,64,192 fwd 12"; do
  set -- $s
  B200SP_LIB_SUFFIX=_tl timeout 120 python tools/tcg2_timeline.py $1 $2 $3 > $O/tl_$2_$(echo $1 | tr , x).txt 2>&1
done
timeout 900 python -m pytest -q tests/test_cli_gpu.py tests/test_next_rows_gpu.py tests/test_krn_gpu.py tests/test_ddp_gpu.py tests/test_optim_ckpt_gpu.py -s 2>&1 | tail -30 > $O/t_tests.log
ls -la $O
