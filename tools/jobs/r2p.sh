#!/bin/bash
# round-2 GPU job P (2 GPUs): DDP parity tests, torchrun CLIs, KRN / DANN bench at N=2
O=gpurun_out/r2p; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest -q tests/test_ddp_gpu.py -s 2>&1 | tail -25 > $O/t_ddp.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
mkdir -p /tmp/cli2 && cd /tmp/cli2
timeout 600 $TR --master-port 29621 $GRAFT_REPO_ROOT/train.py --model_name krn --optimizer adamw --batch_size 4 --synthetic_data 3 --max_epochs 1 --savedir ck --logdir lg --start_over > $GRAFT_REPO_ROOT/$O/cli_train2.log 2>&1; echo "train rc=$?" >> $GRAFT_REPO_ROOT/$O/cli_train2.log; ls ck >> $GRAFT_REPO_ROOT/$O/cli_train2.log
timeout 600 $TR --master-port 29622 $GRAFT_REPO_ROOT/adapt.py --perform_dann --model_name krn --optimizer adamw --batch_size 4 --synthetic_data 2 --max_epochs 1 --savedir ckd --logdir lgd --start_over > $GRAFT_REPO_ROOT/$O/cli_adapt2.log 2>&1; echo "adapt rc=$?" >> $GRAFT_REPO_ROOT/$O/cli_adapt2.log; ls ckd >> $GRAFT_REPO_ROOT/$O/cli_adapt2.log
cd $GRAFT_REPO_ROOT
timeout 600 $TR --master-port 29623 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_krn_n2.json 2> $O/bench_krn_n2.err
B200SP_GRAPH_NCCL=0 timeout 600 $TR --master-port 29625 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_krn_n2_eagernccl.json 2> $O/bench_krn_n2_eagernccl.err
timeout 600 $TR --master-port 29624 bench.py --gpus 2 --steps 50 --warmup 5 --workload dann > $O/bench_dann_n2.json 2> $O/bench_dann_n2.err
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --workload dann --no-cpu-baseline > $O/bench_dann_n1.json 2> $O/bench_dann_n1.err
