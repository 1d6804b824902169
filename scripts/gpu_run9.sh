#!/bin/bash
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run9.log; : > $L
echo "=== repro" >> $L
timeout 120 python scripts/repro_dw.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== pytest gpu (single), no -x" >> $L
timeout 1200 python -m pytest tests -m "gpu and not multigpu" -q >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$" $L | tail -c 7000
