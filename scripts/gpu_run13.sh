#!/bin/bash
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run13.log; : > $L
echo "=== pytest persistent + 2cta" >> $L
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "persistent or 2cta" >> $L 2>&1; echo "exit=$?" >> $L
echo "=== pytest rest" >> $L
timeout 900 python -m pytest tests -m "gpu and not multigpu" -q -k "not persistent and not 2cta" >> $L 2>&1; echo "exit=$?" >> $L
echo "=== perf2" >> $L
timeout 300 python scripts/gemm_case.py perf2 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused (persistent step)" >> $L
timeout 300 python bench.py --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench fused (per-GEMM launches)" >> $L
timeout 300 python bench.py --steps 30 --warmup 5 --no-fused-step >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$" $L | tail -c 8000
