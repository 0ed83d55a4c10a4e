#!/bin/bash
# round-2 GPU job S: stacked-B MMA (2 issues per k-step), with / without converter groups
O=gpurun_out/r2s; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 600 python -m pytest -q tests/test_kernels_gpu.py -k "pw_" 2>&1 | tail -6 > $O/t_kernels.log
B200SP_TCG2_GROUPS=2 timeout 600 python -m pytest -q tests/test_kernels_gpu.py -k "pw_fwd or pw_dgrad" 2>&1 | tail -4 > $O/t_kernels_g2.log
timeout 400 python tools/gemm_bench.py --reps 5 > $O/gemm_bench_stack.txt 2>&1
B200SP_TCG2_GROUPS=2 timeout 400 python tools/gemm_bench.py --reps 5 --ops fwd,dgrad > $O/gemm_bench_stack_g2.txt 2>&1
B200SP_TCG2_STACKB=0 timeout 400 python tools/gemm_bench.py --reps 5 > $O/gemm_bench_nostack.txt 2>&1
timeout 400 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof.txt > $O/bench.json 2> $O/bench.err
B200SP_TCG2_GROUPS=2 timeout 400 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_g2.json 2> $O/bench_g2.err
for s in "9408,96,576 fwd 12" "2352,1024,1280 fwd 12"; do
  set -- $s
  B200SP_LIB_SUFFIX=_tl timeout 120 python tools/tcg2_timeline.py $1 $2 $3 > $O/tl_$2_$(echo $1 | tr , x).txt 2>&1
done
