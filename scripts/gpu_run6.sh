#!/bin/bash
# 2-GPU: VMM-over-sockets + multicast probe, protocol checks, benches
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run6.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
echo "=== symm probe" >> $L
timeout 300 $TR --master-port 29511 scripts/symm_probe.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== multi gpu check" >> $L
timeout 600 $TR --master-port 29514 scripts/multi_gpu_check.py fused two_shot generic >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused N=2" >> $L
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench nccl N=2" >> $L
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --impl nccl --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$|\*\*\*\*|OMP_NUM" $L | tail -c 9000
