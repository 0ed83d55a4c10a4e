#!/bin/bash
# round-2 GPU job H: converter groups (several k-blocks in conversion at once)
O=gpurun_out/r2h; mkdir -p $O
export B200SP_NO_AUTOBUILD=1 B200SP_TCG2=1
timeout 600 python -m pytest -q tests/test_kernels_gpu.py -k "pw_fwd or pw_dgrad or pw_wgrad" 2>&1 | tail -30 > $O/t_kernels.log
timeout 600 python tools/gemm_bench.py --reps 5 > $O/gemm_bench_g4.txt 2>&1
B200SP_TCG2_GROUPS=2 timeout 600 python tools/gemm_bench.py --reps 5 > $O/gemm_bench_g2.txt 2>&1
for s in "2352,1024,1280 fwd 16" "9408,96,576 fwd 20" "9408,64,192 fwd 8"; do
  set -- $s
  B200SP_LIB_SUFFIX=_tl timeout 120 python tools/tcg2_timeline.py $1 $2 $3 > $O/tl_$2_$(echo $1 | tr , x).txt 2>&1
done
timeout 600 python -m pytest -q -x tests/test_krn_gpu.py tests/test_dann_gpu.py tests/test_cli_gpu.py 2>&1 | tail -15 > $O/t_models.log
timeout 600 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof.txt > $O/bench.json 2> $O/bench.err
