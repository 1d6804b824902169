#!/bin/bash
# 1 GPU: last check of the committed tree (defaults only)
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run29.log; : > $L
echo "=== smoke" >> $L
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 >> $L; echo "exit=$?" >> $L
echo "=== pytest -m gpu" >> $L
timeout 400 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 >> $L; echo "exit=$?" >> $L
echo "=== bench" >> $L
timeout 120 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
tail -c 3000 $L | cut -c1-1500
