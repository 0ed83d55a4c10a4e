#!/bin/bash
O=gpurun_out/r3z2; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
for i in 1 2; do timeout 900 python bench.py --no-secondary --no-cpu-baseline 2>> $O/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'])"; done
