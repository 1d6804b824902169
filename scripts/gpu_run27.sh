#!/bin/bash
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run27.log; : > $L
echo "=== pytest engines + persistent (input pipeline on)" >> $L
BFLC_INPUT_PIPELINE=1 timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_kernels.py -q -x -k "engine or fused or checkpoint or persistent" 2>&1 | tail -4 >> $L
echo "=== bench pipeline on" >> $L
BFLC_INPUT_PIPELINE=1 timeout 200 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench pipeline off" >> $L
BFLC_INPUT_PIPELINE=0 timeout 200 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
tail -c 4000 $L | cut -c1-1600
