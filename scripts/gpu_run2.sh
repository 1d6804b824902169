#!/bin/bash
# 1-GPU: pytest -m gpu, smoke, fused bench, NCCL-baseline bench, launch list via ncu.
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run2.log; : > $L
echo "=== pytest" >> $L
timeout 600 python -m pytest tests -m gpu -x -q >> $L 2>&1; echo "exit=$?" >> $L
echo "=== smoke" >> $L
timeout 300 python __graft_entry__.py smoke >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused" >> $L
timeout 300 python bench.py --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench nccl" >> $L
timeout 300 python bench.py --impl nccl --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench reference" >> $L
timeout 60 python bench.py --impl reference >> $L 2>&1; echo "exit=$?" >> $L
echo "=== ncu launch list (fused, no graph)" >> $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fused.csv python bench.py --steps 2 --warmup 3 --no-graph > gpurun_out/ncu_b.log 2>&1; echo "exit=$?" >> $L
tail -c 6000 $L
