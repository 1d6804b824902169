#!/bin/bash
# compute-sanitizer memcheck over single-GPU kernel cases (bounded: small shapes, short timeouts)
mkdir -p gpurun_out
export BFLC_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
L=gpurun_out/sanitize.log; : > $L
for c in kk_300_200_784 epi xent; do
  echo "=== memcheck $c" >> $L
  timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/gemm_case.py $c 2>&1 | grep -E "RESULT|ERROR SUMMARY|Error|error:|Invalid" | cut -c1-300 | head -8 >> $L
  echo "exit=${PIPESTATUS[0]}" >> $L
done
tail -c 3000 $L
