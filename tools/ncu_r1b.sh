set -x
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/r1b_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tcgemm -s 3 -c 1 -o gpurun_out/r1b_pwfwd_2352x1024x1280 python tools/gemm_bench.py --shapes 2352,1024,1280 --ops fwd --reps 1 > gpurun_out/n1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tcgemm -s 3 -c 1 -o gpurun_out/r1b_pwdgrad_602112x16x32 python tools/gemm_bench.py --shapes 602112,16,32 --ops dgrad --reps 1 > gpurun_out/n2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tcgemm -s 3 -c 1 -o gpurun_out/r1b_pwwgrad_2352x1024x1280 python tools/gemm_bench.py --shapes 2352,1024,1280 --ops wgrad --reps 1 > gpurun_out/n3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dwr_bwd -s 18 -c 1 -o gpurun_out/r1b_dwbwd_96s2 python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline > gpurun_out/n4.log 2>&1
python tools/gemm_bench.py --reps 5 > gpurun_out/r1b_gemm_bench.txt 2>&1
ls -la gpurun_out
