# round-1 final evidence (final code state of the round)
# row-decomposed style conv, split-K SPN FC)
set -x
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1i_smoke.log 2>&1
timeout 280 python bench.py --steps 30 --warmup 5 --profile-out gpurun_out/r1i_step_profile.txt > gpurun_out/r1i_bench.json 2> gpurun_out/r1i_bench.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1i_bench_reference.json 2> gpurun_out/r1i_bench_reference.err
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1i_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-secondary > gpurun_out/r1i_ncu_bench.log 2>&1
timeout 60 python tools/spn_bench.py > gpurun_out/r1i_spn_profile.txt 2>&1
timeout 60 python tools/styleaug_bench.py > gpurun_out/r1i_styleaug_profile.txt 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:convtc -s 51 -c 1 -o gpurun_out/r1i_convtc_res3x3_128 python tools/styleaug_bench.py --reps 1 > gpurun_out/n5.log 2>&1
tail -1 gpurun_out/r1i_smoke.log
