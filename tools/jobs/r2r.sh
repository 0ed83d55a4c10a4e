#!/bin/bash
# 2 GPUs: DDP parity tests + torchrun train.py with the default (eager all-reduce between the two graphs)
O=gpurun_out/r2r; mkdir -p $O
export B200SP_NO_AUTOBUILD=1
timeout 420 python -m pytest -q tests/test_ddp_gpu.py -s 2>&1 | tail -12 > $O/t_ddp.log
mkdir -p /tmp/cli2 && cd /tmp/cli2
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 $GRAFT_REPO_ROOT/train.py --model_name krn --optimizer adamw --batch_size 4 --synthetic_data 3 --max_epochs 1 --savedir ck --logdir lg --start_over > $GRAFT_REPO_ROOT/$O/cli_train2.log 2>&1; echo "train rc=$?" >> $GRAFT_REPO_ROOT/$O/cli_train2.log; ls ck >> $GRAFT_REPO_ROOT/$O/cli_train2.log
