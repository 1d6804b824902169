#!/bin/bash
# 1-GPU: all gpu tests, phases, kernel bench, perf, bench (PDL on/off)
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run8.log; : > $L
echo "=== pytest gpu (single)" >> $L
timeout 900 python -m pytest tests -m "gpu and not multigpu" -x -q >> $L 2>&1; echo "exit=$?" >> $L
echo "=== gemm phases" >> $L
timeout 120 python scripts/gemm_phases.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== kernel bench B=512" >> $L
timeout 300 python scripts/kernel_bench.py 512 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== perf" >> $L
timeout 300 python scripts/gemm_case.py perf >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused" >> $L
timeout 300 python bench.py --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused no-pdl" >> $L
BFLC_PDL=0 timeout 300 python bench.py --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$" $L | tail -c 9000
