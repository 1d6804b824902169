#!/bin/bash
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run26.log; : > $L
echo "=== pytest persistent + engines" >> $L
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q -x -k "persistent or engine or fused or checkpoint" 2>&1 | tail -4 >> $L
echo "=== phases" >> $L
timeout 100 python scripts/mlp_phases.py 2>&1 | grep -E "PHASES|Error|error" >> $L
echo "=== bench" >> $L
timeout 200 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
tail -c 3500 $L | cut -c1-1200
