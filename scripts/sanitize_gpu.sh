#!/bin/bash
# compute-sanitizer (memcheck + racecheck on shared memory) over single-GPU kernel cases
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/sanitize.log; : > $L
for tool in memcheck racecheck; do
  for c in kk_300_200_784 epi xent elem; do
    echo "=== $tool $c" >> $L
    timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python scripts/gemm_case.py $c 2>&1 | grep -E "RESULT|ERROR SUMMARY|Error|error:" | head -8 >> $L
    echo "exit=${PIPESTATUS[0]}" >> $L
  done
done
tail -c 4000 $L
