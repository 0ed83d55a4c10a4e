#!/bin/bash
O=gpurun_out/r2v; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 300 python -m pytest -q tests/test_kernels_gpu.py -k "dw" 2>&1 | tail -6 > $O/t_dw.log
timeout 300 python -m pytest -q -x tests/test_krn_gpu.py tests/test_dann_gpu.py 2>&1 | tail -4 > $O/t_models.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof.txt > $O/bench.json 2> $O/bench.err
B200SP_DW_SMALLSEG=0 timeout 300 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof_old.txt > $O/bench_old.json 2> $O/bench_old.err
