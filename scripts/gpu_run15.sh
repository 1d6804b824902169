#!/bin/bash
# 1 GPU: block-scaled fp8 bring-up + ncu --set full captures of the hot kernels + launch list
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/run15.log; : > $L
echo "=== pytest mx8 + generic lenet" >> $L
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_engine.py -q -x -k "mx8 or lenet" >> $L 2>&1; echo "exit=$?" >> $L
echo "=== perf3" >> $L
timeout 200 python scripts/gemm_case.py perf3 >> $L 2>&1; echo "exit=$?" >> $L
NCU="ncu --set full --clock-control none --import-source on"
for c in ncu_2cta ncu_1cta ncu_mx8; do
  echo "=== $c" >> $L
  timeout 200 $NCU -k regex:gemm -s 2 -c 1 -f -o gpurun_out/$c python scripts/gemm_case.py $c >> $L 2>&1; echo "exit=$?" >> $L
done
echo "=== ncu mlp_round" >> $L
timeout 300 $NCU -k regex:mlp_round -s 3 -c 1 -f -o gpurun_out/ncu_mlp_round python bench.py --no-graph --steps 3 --warmup 3 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== ncu consensus/upload" >> $L
timeout 300 $NCU -k regex:"k_consensus|k_upload" -s 6 -c 2 -f -o gpurun_out/ncu_fed python bench.py --no-graph --steps 3 --warmup 3 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== launch list" >> $L
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --no-graph --steps 3 --warmup 3 >> $L 2>&1; echo "exit=$?" >> $L
echo "=== bench (1 GPU, final code)" >> $L
timeout 200 python bench.py --steps 40 --warmup 5 >> $L 2>&1; echo "exit=$?" >> $L
grep -vE "Warn|warn|^$|\*\*\*\*|OMP_NUM|==PROF==" $L | tail -c 7000
