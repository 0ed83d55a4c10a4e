# round-1 final refresh (code state: measured GEMM dispatch, 4-deep TMEM ring, async wgrad)
set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1f_smoke.log 2>&1
timeout 280 python bench.py --steps 30 --warmup 5 --profile-out gpurun_out/r1f_step_profile.txt > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1f_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-secondary > gpurun_out/r1f_ncu_bench.log 2>&1
timeout 60 python tools/spn_bench.py > gpurun_out/r1f_spn_profile.txt 2>&1
timeout 60 python tools/styleaug_bench.py > gpurun_out/r1f_styleaug_profile.txt 2>&1
tail -2 gpurun_out/r1f_smoke.log
