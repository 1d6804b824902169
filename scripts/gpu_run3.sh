#!/bin/bash
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run3.log; : > $L
echo "=== kernel bench B=512" >> $L
timeout 300 python scripts/kernel_bench.py 512 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== kernel bench B=4096" >> $L
timeout 300 python scripts/kernel_bench.py 4096 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== ncu full: gemm small + big" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 30 -c 4 -f -o gpurun_out/prof_gemm_small python scripts/gemm_case.py perf_small > gpurun_out/ncu_small.log 2>&1; echo "exit=$?" >> $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -f -o gpurun_out/prof_gemm_big python scripts/gemm_case.py perf_big > gpurun_out/ncu_big.log 2>&1; echo "exit=$?" >> $L
echo "=== pytest" >> $L
timeout 600 python -m pytest tests -m gpu -x -q >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench nccl" >> $L
timeout 300 python bench.py --impl nccl --steps 30 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
tail -c 5000 $L
