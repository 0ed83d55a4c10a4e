# round-1 final evidence run: bench lines, ncu launch list, ncu --set full captures (each step under its own timeout)
set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1d_smoke.log 2>&1
timeout 280 python bench.py --steps 30 --warmup 5 --profile-out gpurun_out/r1d_step_profile.txt > gpurun_out/r1d_bench.json 2> gpurun_out/r1d_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1d_bench_reference.json 2> gpurun_out/r1d_bench_reference.err
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-secondary > gpurun_out/r1d_ncu_bench.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:dwr_bwd -s 18 -c 1 -o gpurun_out/r1d_dwbwd_96s2 python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-secondary > gpurun_out/n4.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:convtc -s 51 -c 1 -o gpurun_out/r1d_convtc_res3x3_128 python tools/styleaug_bench.py --reps 1 > gpurun_out/n5.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:adamw_kernel -s 2 -c 1 -o gpurun_out/r1d_adamw python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-secondary > gpurun_out/n6.log 2>&1
timeout 100 python tools/gemm_bench.py --reps 5 > gpurun_out/r1d_gemm_bench.txt 2>&1
ls -la gpurun_out | tail -20
