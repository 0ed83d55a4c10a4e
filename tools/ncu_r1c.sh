ncu --set full --clock-control none --import-source on -k regex:tcgemm -s 3 -c 1 -o gpurun_out/r1c_pwfwd_602112x16x32 python tools/gemm_bench.py --shapes 602112,16,32 --ops fwd --reps 1 > gpurun_out/n1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tcgemm -s 3 -c 1 -o gpurun_out/r1c_pwwgrad_602112x16x32 python tools/gemm_bench.py --shapes 602112,16,32 --ops wgrad --reps 1 > gpurun_out/n2.log 2>&1
ls -la gpurun_out/*.ncu-rep
