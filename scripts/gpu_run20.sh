#!/bin/bash
# one 8-GPU box: FedAvg as two-shot (reduce own slice + multicast publish) for the small model
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run20.log; : > $L
tr() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port $2"; }
echo "=== bench fused N=8 two-shot" >> $L
timeout 150 $(tr 8 29516) bench.py --gpus 8 --steps 30 --warmup 5 --two-shot on >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused N=4 two-shot" >> $L
timeout 150 $(tr 4 29512) bench.py --gpus 4 --steps 30 --warmup 5 --two-shot on >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$|\*\*\*\*|OMP_NUM" $L | cut -c1-2500 | tail -c 6000
