#!/bin/bash
# round-2 final evidence (1 GPU): GPU test suite, smoke, bench (both arms), ncu launch lists, ncu --set full of the top kernels
O=gpurun_out/r2_final; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 1200 python -m pytest -q tests -m gpu -s 2>&1 | tail -40 > $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 1500 python bench.py --profile-out $O/step_profile.txt > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py --workload dann --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_dann.json 2> $O/bench_dann.err
# every launch of two eager + graph steps with its device time (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline > $O/ncu_bench.log 2>&1
# the reference's own CUDA path under the same launch-list pass (3 iterations of its loop)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches_reference_cuda.csv python -c "
import sys; sys.path.insert(0, 'baseline')
import ref_runner as R
print(R.time_krn_train('cuda:0', warmup=1, steps=2))" > $O/ncu_ref.log 2>&1
N="timeout 200 ncu --set full --clock-control none --import-source on -s 3 -c 1"
$N -k regex:tcgemm2 -o $O/full_gemm_fwd_2352x1024x1280 python tools/gemm_bench.py --shapes 2352,1024,1280 --ops fwd --reps 1 > $O/n1.log 2>&1
$N -k regex:tcgemm2 -o $O/full_gemm_dgrad_602112x96x16 python tools/gemm_bench.py --shapes 602112,96,16 --ops dgrad --reps 1 > $O/n2.log 2>&1
ls -la $O
