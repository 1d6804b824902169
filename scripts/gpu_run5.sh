#!/bin/bash
# 1-GPU: new staged epilogue -> numerics, phases, kernel bench, model tests, bench
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run5.log; : > $L
echo "=== pytest kernels+engine" >> $L
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -x -q >> $L 2>&1; echo "exit=$?" >> $L
echo "=== gemm phases" >> $L
timeout 120 python scripts/gemm_phases.py >> $L 2>&1; echo "exit=$?" >> $L
echo "=== kernel bench B=512" >> $L
timeout 300 python scripts/kernel_bench.py 512 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== perf" >> $L
timeout 300 python scripts/gemm_case.py perf >> $L 2>&1; echo "exit=$?" >> $L
echo "=== pytest models" >> $L
timeout 900 python -m pytest tests/test_gpu_models.py -q >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused" >> $L
timeout 300 python bench.py --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$" $L | tail -c 9000
