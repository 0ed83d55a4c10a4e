set -x
timeout 150 ncu --set full --clock-control none --import-source on -k regex:tcgemm -s 3 -c 1 -o gpurun_out/r1j_pwfwd_2352x1024x1280 python tools/gemm_bench.py --shapes 2352,1024,1280 --ops fwd --reps 1 > gpurun_out/n1.log 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:tcgemm -s 3 -c 1 -o gpurun_out/r1j_pwdgrad_602112x96x16 python tools/gemm_bench.py --shapes 602112,96,16 --ops dgrad --reps 1 > gpurun_out/n2.log 2>&1
ls gpurun_out/r1j*
