#!/bin/bash
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run25.log; : > $L
echo "=== phases" >> $L
timeout 100 python scripts/mlp_phases.py 2>&1 | grep -E "PHASES|Error|error" >> $L
bash scripts/sanitize_gpu.sh > /dev/null 2>&1
cat gpurun_out/sanitize.log >> $L
tail -c 3000 $L
