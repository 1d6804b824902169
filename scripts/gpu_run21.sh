#!/bin/bash
# 1 GPU: cluster plan (chain 4: DSMEM-assembled h tile, cluster barrier) vs chain 3
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run21.log; : > $L
for c in 4 3; do
  echo "=== chain=$c" >> $L
  BFLC_MLP_CHAIN=$c timeout 200 python -m pytest tests/test_gpu_kernels.py -q -x -k "persistent" 2>&1 | tail -3 >> $L
  BFLC_MLP_CHAIN=$c timeout 100 python scripts/mlp_phases.py 2>&1 | grep -E "PHASES|Error|error" >> $L
  echo "--- bench chain=$c" >> $L
  BFLC_MLP_CHAIN=$c timeout 200 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
done
echo "=== engine tests chain=4" >> $L
BFLC_MLP_CHAIN=4 timeout 300 python -m pytest tests/test_gpu_engine.py -q 2>&1 | tail -3 >> $L
grep -vE "Warn|warn|^$|\*\*\*\*|OMP_NUM|==PROF==" $L | cut -c1-1300 | tail -c 7000
