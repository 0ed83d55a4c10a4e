#!/bin/bash
O=gpurun_out/r3t; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_krn_gpu.py -m gpu -x -q -k "bn_apply or train or step" 2>&1 | tail -2
timeout 900 python bench.py --steps 50 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/step_profile.txt 2> $O/bench.err | head -c 330
echo; grep -E "bn_apply|dw_fwd|dw_bwd|stem|head|reorg" $O/step_profile.txt | head -80
