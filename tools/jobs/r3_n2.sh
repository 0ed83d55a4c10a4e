#!/bin/bash
# 2 GPUs at the final code state: DDP parity tests, KRN / DANN bench lines at N=2
O=gpurun_out/r3_n2; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 420 python -m pytest -q tests/test_ddp_gpu.py 2>&1 | tail -5 > $O/t_ddp.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 50 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_krn_n2.json 2> $O/bench_krn_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --workload dann --steps 30 --warmup 5 --no-secondary --no-cpu-baseline > $O/bench_dann_n2.json 2> $O/bench_dann_n2.err
cat $O/t_ddp.log; head -c 420 $O/bench_krn_n2.json; echo; head -c 300 $O/bench_dann_n2.json
