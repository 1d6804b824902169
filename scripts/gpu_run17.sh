#!/bin/bash
# 1 GPU: phase-plan variants of the persistent trainer: correctness + in-kernel stamps + bench
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run17.log; : > $L
for c in 3 1 0; do for e in 1 0; do
  echo "=== chain=$c epiopt=$e" >> $L
  BFLC_MLP_CHAIN=$c BFLC_MLP_EPIOPT=$e timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -k "persistent" 2>&1 | tail -2 >> $L
  BFLC_MLP_CHAIN=$c BFLC_MLP_EPIOPT=$e timeout 100 python scripts/mlp_phases.py 2>&1 | grep -E "PHASES|Error|error" >> $L
done; done
echo "=== bench default (chain 3, epiopt 1)" >> $L
timeout 200 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench chain 0 epiopt 1" >> $L
BFLC_MLP_CHAIN=0 timeout 200 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench chain 0 epiopt 0" >> $L
BFLC_MLP_CHAIN=0 BFLC_MLP_EPIOPT=0 timeout 200 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$|\*\*\*\*|OMP_NUM|==PROF==" $L | cut -c1-1200 | tail -c 9000
