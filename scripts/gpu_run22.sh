#!/bin/bash
# 1 GPU: cluster plan test (run21) followed by ncu captures + full single-GPU pytest (run19)
bash scripts/gpu_run21.sh > /dev/null 2>&1
bash scripts/gpu_run19.sh > /dev/null 2>&1
tail -c 3000 gpurun_out/run21.log; tail -c 2500 gpurun_out/run19.log
