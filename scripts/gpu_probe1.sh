#!/bin/bash
# First GPU bring-up: GEMM descriptor variants + epilogues + perf. Each case in its own
# process under `timeout` so a trap/hang cannot take the box down.
mkdir -p gpurun_out
LOG=gpurun_out/probe1.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $LOG 2>&1
export BFLC_NO_AUTOBUILD=1
export PYTHONPATH=$PWD:$PYTHONPATH
run() { echo "=== $1" >> $LOG; timeout 120 python scripts/gemm_case.py $1 >> $LOG 2>&1; echo "exit=$?" >> $LOG; }
run kk_128_64_64
run kk_256_256_512
run kk_200_62_784
run kk_1000_300_1000
run kmn_256_256_512
run kmn_256_256_512_1024_8192
run kmn_200_256_62
run mnmn_256_256_512
run mnmn_256_256_512_1024_8192
run mnmn_62_256_1000
run fp8_256_256_512
run epi
run xent
run elem
run perf
grep -E "^===|RESULT|exit=|Error|error|rror:" $LOG | tail -80
