#!/bin/bash
# 2 GPUs: e2e input pipeline on trainer + committee ranks
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run28.log; : > $L
BFLC_INPUT_PIPELINE=1 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$|\*\*\*\*|OMP_NUM" $L | cut -c1-2500 | tail -c 4000
