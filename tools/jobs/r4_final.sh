#!/bin/bash
# closing evidence of the final commit (1 GPU): GPU test suite, smoke, bench (both arms), DANN bench, step profile
O=gpurun_out/r4_final; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 1500 python -m pytest -q tests -m gpu 2>&1 | tail -15 > $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
timeout 1500 python bench.py --profile-out $O/step_profile.txt > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py --workload dann --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_dann.json 2> $O/bench_dann.err
tail -3 $O/pytest_gpu.log; tail -2 $O/smoke.log; head -c 700 $O/bench.json
