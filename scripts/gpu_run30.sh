#!/bin/bash
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH BFLC_MLP_EXPERIMENTAL=1
timeout 60 python scripts/plan4_diag.py > gpurun_out/run30.log 2>&1
tail -c 1500 gpurun_out/run30.log
