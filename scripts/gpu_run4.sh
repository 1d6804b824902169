#!/bin/bash
# 2-GPU: phases (rank-less), substrate probe, fused + nccl bench at N=2
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run4.log; : > $L
echo "=== gemm phases" >> $L
timeout 120 python scripts/gemm_phases.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== symm probe" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29511 scripts/symm_probe.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused N=2" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench nccl N=2" >> $L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --impl nccl --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$|\*\*\*\*" $L | tail -c 7000
