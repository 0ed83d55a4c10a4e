#!/bin/bash
# round-2 GPU job A: validate the r2 candidates (dw v2, direct FFMA pointwise, lean tf32 split) and time the
# reference's own CUDA path (cuDNN) on the same box.  Everything lands in gpurun_out/r2a/.
O=gpurun_out/r2a; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/smi.txt
T="timeout 900 python -m pytest -q -x"
$T tests/test_kernels_gpu.py 2>&1 | tail -8 > $O/t0_kernels_default.log
B200SP_DW=2 $T tests/test_kernels_gpu.py -k dw 2>&1 | tail -15 > $O/t1_dw2_kernels.log
B200SP_DW=2 $T tests/test_krn_gpu.py tests/test_dann_gpu.py 2>&1 | tail -15 > $O/t1_dw2_models.log
B200SP_PWDIRECT=1 timeout 900 python -m pytest -q tests/test_kernels_gpu.py -k "pw_fwd or pw_dgrad" 2>&1 | tail -40 > $O/t2_pwdirect_kernels.log
B200SP_PWDIRECT=1 $T tests/test_krn_gpu.py 2>&1 | tail -15 > $O/t2_pwdirect_models.log
B200SP_LIB_SUFFIX=_lean timeout 900 python -m pytest -q tests/test_kernels_gpu.py tests/test_tc_gpu.py -k "pw_ or gemm or tc" 2>&1 | tail -15 > $O/t3_lean_kernels.log
i=0
for cfg in "X=0" "B200SP_DW=2" "B200SP_PWDIRECT=1" "B200SP_DW=2 B200SP_PWDIRECT=1" "B200SP_LIB_SUFFIX=_lean" "B200SP_LIB_SUFFIX=_lean B200SP_DW=2 B200SP_PWDIRECT=1"; do
  env $cfg timeout 600 python bench.py --steps 30 --warmup 5 --no-secondary --no-cpu-baseline --profile-out $O/prof_$i.txt > $O/bench_$i.json 2> $O/bench_$i.err
  echo "$i: $cfg" >> $O/bench_index.txt
  i=$((i+1))
done
B200SP_PWDIRECT=1 timeout 600 python tools/gemm_bench.py --reps 5 --ops fwd,dgrad --shapes 602112,16,32 602112,96,16 150528,24,96 150528,144,24 150528,24,144 37632,32,144 37632,192,32 37632,32,192 > $O/gemm_bench_pwdirect.txt 2>&1
# the reference's own CUDA path (unmodified, baseline/_ref): fp32 and --use_fp16 autocast, 20 warm-up + 100 timed steps
timeout 900 python - > $O/reference_cuda.json 2> $O/reference_cuda.err <<'PY'
import json, sys
sys.path.insert(0, '.')
from baseline import ref_runner as R
out = {}
out['krn_fp32'] = R.time_krn_train('cuda:0', warmup=20, steps=100)
out['krn_amp'] = R.time_krn_train('cuda:0', warmup=20, steps=100, fp16=True)
out['krn_fp32_styleaug'] = R.time_krn_train('cuda:0', warmup=10, steps=50, style=True)
out['dann_fp32'] = R.time_dann_train('cuda:0', warmup=10, steps=50)
out['styleaug_fwd'] = R.time_styleaug('cuda:0', warmup=5, steps=30)
print(json.dumps(out, indent=1))
PY
ls -la $O
