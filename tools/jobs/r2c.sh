#!/bin/bash
# round-2 GPU job C: ncu --set full captures of the second-generation GEMM (three representative launches) + new tests
O=gpurun_out/r2c; mkdir -p $O
export B200SP_NO_AUTOBUILD=1 B200SP_TCG2=1
N="timeout 200 ncu --set full --clock-control none --import-source on -k regex:tcgemm2 -s 3 -c 1"
$N -o $O/fwd_2352x1024x1280 python tools/gemm_bench.py --shapes 2352,1024,1280 --ops fwd --reps 1 > $O/n1.log 2>&1
$N -o $O/fwd_602112x96x16 python tools/gemm_bench.py --shapes 602112,96,16 --ops fwd --reps 1 > $O/n2.log 2>&1
$N -o $O/dgrad_150528x24x144 python tools/gemm_bench.py --shapes 150528,24,144 --ops dgrad --reps 1 > $O/n3.log 2>&1
$N -o $O/fwd_9408x96x576 python tools/gemm_bench.py --shapes 9408,96,576 --ops fwd --reps 1 > $O/n4.log 2>&1
$N -o $O/wgrad_602112x16x32 python tools/gemm_bench.py --shapes 602112,16,32 --ops wgrad --reps 1 > $O/n5.log 2>&1
timeout 600 python -m pytest -q tests/test_optim_ckpt_gpu.py tests/test_styleaug_gpu.py -s 2>&1 | tail -40 > $O/t_new_tests.log
ls -la $O
