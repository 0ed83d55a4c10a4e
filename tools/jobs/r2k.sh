#!/bin/bash
# round-2 GPU job K: trimmed wgrad A tiles (+groups), PDL on depthwise / bn_apply, tf32 tests
O=gpurun_out/r2k; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 900 python -m pytest -q tests/test_kernels_gpu.py tests/test_krn_tf32_gpu.py tests/test_krn_gpu.py tests/test_dann_gpu.py -s 2>&1 | tail -25 > $O/t_tests.log
for g in 1 2 4; do B200SP_TCG2_WGRAD_GROUPS=$g timeout 600 python tools/gemm_bench.py --reps 5 --ops wgrad > $O/gemm_bench_wgrad_g$g.txt 2>&1; done
timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof.txt > $O/bench.json 2> $O/bench.err
B200SP_TCG2_WGRAD_GROUPS=4 timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_wg4.json 2> $O/bench_wg4.err
B200SP_TCG2_WGRAD=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_r1wgrad.json 2> $O/bench_r1wgrad.err
B200SP_PDL=0 timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_pdl0.json 2> $O/bench_pdl0.err
