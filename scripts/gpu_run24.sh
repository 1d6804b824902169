#!/bin/bash
# 1 GPU: final tree -- graft smoke, full gpu pytest, default bench
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run24.log; : > $L
echo "=== smoke" >> $L
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 >> $L; echo "exit=$?" >> $L
echo "=== pytest -m gpu" >> $L
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 >> $L; echo "exit=$?" >> $L
echo "=== bench" >> $L
timeout 200 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench reference arm" >> $L
timeout 60 python bench.py --impl reference >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$|\*\*\*\*|OMP_NUM" $L | cut -c1-1500 | tail -c 5000
