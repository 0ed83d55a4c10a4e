#!/bin/bash
# round-2 GPU job N: stem_wgrad2 (register tile), dw occupancy variant, full suite, secondaries
O=gpurun_out/r2n; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 1200 python -m pytest -q tests -m gpu 2>&1 | tail -15 > $O/t_all_gpu.log
timeout 900 python bench.py --steps 50 --warmup 5 --no-reference-cuda --no-cpu-baseline --profile-out $O/prof.txt > $O/bench.json 2> $O/bench.err
B200SP_LIB_SUFFIX=_occ4 timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof_occ4.txt > $O/bench_occ4.json 2> $O/bench_occ4.err
B200SP_TCG2_WGRAD=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_r1wgrad.json 2> $O/bench_r1wgrad.err
