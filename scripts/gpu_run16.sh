#!/bin/bash
# 1 GPU: fused chain (fwd1->xent->dh) + epilogue optimizer in the persistent training kernel,
# one-launch validation chain
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run16.log; : > $L
echo "=== pytest persistent (chain + epilogue optimizer)" >> $L
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "persistent" >> $L 2>&1; echo "exit=$?" >> $L
echo "=== pytest persistent (chain, flat optimizer phase)" >> $L
BFLC_MLP_EPIOPT=0 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "persistent" >> $L 2>&1; echo "exit=$?" >> $L
echo "=== pytest persistent (chain off)" >> $L
BFLC_MLP_CHAIN=0 timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "persistent" >> $L 2>&1; echo "exit=$?" >> $L
echo "=== pytest engines (val chain) + mx8 models" >> $L
timeout 400 python -m pytest tests/test_gpu_engine.py tests/test_gpu_models.py -q -k "engine or mx8 or fused or checkpoint" >> $L 2>&1; echo "exit=$?" >> $L
echo "=== pytest engines (val chain off)" >> $L
BFLC_VAL_CHAIN=0 timeout 400 python -m pytest tests/test_gpu_engine.py -q -k "fused" >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench all on" >> $L
timeout 200 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench val chain off" >> $L
BFLC_VAL_CHAIN=0 timeout 200 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench epilogue optimizer off" >> $L
BFLC_MLP_EPIOPT=0 timeout 200 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$|\*\*\*\*|OMP_NUM|==PROF==" $L | cut -c1-1300 | tail -c 8000
