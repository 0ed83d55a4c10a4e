#!/bin/bash
# round-2 GPU job M: converter loop without divisions; stem_wgrad2 / head_fwd2; group sweep
O=gpurun_out/r2m; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 900 python -m pytest -q tests/test_kernels_gpu.py tests/test_krn_gpu.py tests/test_krn_tf32_gpu.py -s 2>&1 | tail -25 > $O/t_tests.log
for g in 1 2 4; do B200SP_TCG2_GROUPS=$g B200SP_TCG2_WGRAD_GROUPS=$g timeout 600 python tools/gemm_bench.py --reps 5 > $O/gemm_bench_g$g.txt 2>&1; done
timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof.txt > $O/bench.json 2> $O/bench.err
B200SP_TCG2_GROUPS=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_g1.json 2> $O/bench_g1.err
B200SP_TCG2_GROUPS=4 B200SP_TCG2_WGRAD_GROUPS=2 timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_g4w2.json 2> $O/bench_g4w2.err
