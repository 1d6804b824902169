#!/bin/bash
# N-GPU (N from $1, default 8): protocol checks + fused/nccl bench
N=${1:-8}
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run7_n$N.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1"
echo "=== multi gpu check N=$N" >> $L
timeout 420 $TR --master-port 29514 scripts/multi_gpu_check.py fused byzantine two_shot generic >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused N=$N" >> $L
timeout 240 $TR --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench nccl N=$N" >> $L
timeout 240 $TR --master-port 29513 bench.py --gpus $N --impl nccl --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$|\*\*\*\*|OMP_NUM" $L | tail -c 9000
