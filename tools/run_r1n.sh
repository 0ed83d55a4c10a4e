# r1_n: the (f)-row additions after r1_m (vectorised sampler, raw-frame loader, fused optimizers inside the captured step)
set -x
timeout 200 python -m pytest tests/test_next_rows_gpu.py -q > gpurun_out/r1n_next_rows.log 2>&1
tail -12 gpurun_out/r1n_next_rows.log
timeout 60 python tools/inputpipe_bench.py > gpurun_out/r1n_inputpipe.txt 2>&1
cat gpurun_out/r1n_inputpipe.txt
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1n_inputpipe_launches.csv python tools/inputpipe_bench.py --reps 2 > gpurun_out/r1n_ncu.log 2>&1
tail -3 gpurun_out/r1n_ncu.log
