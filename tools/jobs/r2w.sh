#!/bin/bash
O=gpurun_out/r2w; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 300 python tools/spn_bench.py > $O/spn_profile.txt 2>&1
timeout 300 python tools/styleaug_bench.py > $O/styleaug_profile.txt 2>&1
