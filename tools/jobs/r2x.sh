#!/bin/bash
O=gpurun_out/r2x; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 600 python -m pytest tests/test_aux_kernels_gpu.py tests/test_spn_gpu.py -m gpu -x -q > $O/tests.txt 2>&1; echo "tests rc=$?" >> $O/tests.txt
timeout 300 python tools/spn_bench.py > $O/spn_profile_stream.txt 2>&1
B200SP_FC_STREAM=0 timeout 300 python tools/spn_bench.py > $O/spn_profile_splitk.txt 2>&1
tail -5 $O/tests.txt; grep -E "fc_|ms" $O/spn_profile_stream.txt | head -20
