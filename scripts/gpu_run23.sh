#!/bin/bash
# 2 GPUs: the other BASELINE.json model configs through the generic engine
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run23.log; : > $L
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1"
for c in mlp_fp8 lenet5_fp8 resnet18_byz bert; do
  echo "=== $c" >> $L
  timeout 140 $TR --master-port 29541 scripts/bench_models.py --configs $c --rounds 4 >> $L 2>&1; rc=$?; echo "exit=$rc" >> $L
  if [ $rc -eq 124 ]; then echo "timeout -> stop" >> $L; break; fi
done
grep -vE "Warn|warn|^$|\*\*\*\*|OMP_NUM" $L | cut -c1-900 | tail -c 6000
