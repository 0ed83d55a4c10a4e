#!/bin/bash
O=gpurun_out/r3r; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 900 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/step_profile.txt > $O/bench.json 2> $O/bench.err
head -60 $O/step_profile.txt
