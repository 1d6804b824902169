#!/bin/bash
# 1 GPU: ncu source-level capture of the persistent trainer + validation chain + consensus
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run19.log; : > $L
NCU="ncu --set full --clock-control none --import-source on"
echo "=== ncu mlp_round (phases script)" >> $L
timeout 300 $NCU -k regex:mlp_round -s 2 -c 1 -f -o gpurun_out/ncu_mlp_round2 python scripts/mlp_phases.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== ncu mlp_val + consensus + upload (bench, no graph)" >> $L
timeout 300 $NCU -k regex:"mlp_val|k_consensus|k_upload" -s 9 -c 3 -f -o gpurun_out/ncu_fed2 python bench.py --no-graph --steps 3 --warmup 3 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== pytest all gpu (single GPU)" >> $L
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 >> $L; echo "exit=$?" >> $L
echo "=== bench" >> $L
timeout 200 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$|\*\*\*\*|OMP_NUM|==PROF==|==WARNING==" $L | cut -c1-1500 | tail -c 6000
