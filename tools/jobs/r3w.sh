#!/bin/bash
O=gpurun_out/r3w; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 900 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/step_profile.txt 2> $O/bench.err | head -c 330
echo; head -24 $O/step_profile.txt; grep -A200 "every launch" $O/step_profile.txt | awk '{print $1,$2,$8,$9}' | head -200
