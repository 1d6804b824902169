#!/bin/bash
# one 8-GPU box: protocol checks + fused bench at N=2, 4, 8 with the final kernels
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run18.log; : > $L
tr() { echo "python -m torch.distributed.run --nnodes=1 --nproc-per-node=$1 --master-addr 127.0.0.1 --master-port $2"; }
echo "=== multi gpu check N=8 (fused byzantine two_shot generic)" >> $L
timeout 400 $(tr 8 29515) scripts/multi_gpu_check.py fused byzantine two_shot generic >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused N=8" >> $L
timeout 200 $(tr 8 29516) bench.py --gpus 8 --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused N=4" >> $L
timeout 200 $(tr 4 29512) bench.py --gpus 4 --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused N=2" >> $L
timeout 200 $(tr 2 29513) bench.py --gpus 2 --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== multi gpu check N=4 (fused)" >> $L
timeout 300 $(tr 4 29514) scripts/multi_gpu_check.py fused >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$|\*\*\*\*|OMP_NUM" $L | cut -c1-2500 | tail -c 9000
