#!/bin/bash
# one 8-GPU box: protocol checks + fused bench at N=4 and N=8 (nccl arm numbers: profiles/eight_gpu_r1_first.log)
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run14.log; : > $L
tr() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port $2"; }
echo "=== multi gpu check N=4 (fused generic)" >> $L
timeout 300 $(tr 4 29514) scripts/multi_gpu_check.py fused generic >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused N=4" >> $L
timeout 200 $(tr 4 29512) bench.py --gpus 4 --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== multi gpu check N=8 (fused byzantine two_shot)" >> $L
timeout 300 $(tr 8 29515) scripts/multi_gpu_check.py fused byzantine two_shot >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused N=8" >> $L
timeout 200 $(tr 8 29516) bench.py --gpus 8 --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench nccl N=8" >> $L
timeout 200 $(tr 8 29517) bench.py --gpus 8 --impl nccl --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$|\*\*\*\*|OMP_NUM" $L | tail -c 9000
