#!/bin/bash
O=gpurun_out/r3_final2; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 1500 python bench.py --profile-out $O/step_profile.txt > $O/bench.json 2> $O/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_reference.json 2> $O/bench_reference.err
timeout 600 python bench.py --workload dann --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_dann.json 2> $O/bench_dann.err
head -c 900 $O/bench.json
