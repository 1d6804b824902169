#!/bin/bash
# One gpurun call = one "session": runs each quoted command under its own timeout, tees the
# output into gpurun_out/<tag>/NN.log, and prints a one-line verdict per command at the end.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_session.sh <tag> "cmd 1" "cmd 2" ...'
# (per-command limit: BFLC_CMD_TIMEOUT seconds, default 600)
tag=$1; shift
out=gpurun_out/$tag
mkdir -p "$out"
export PYTHONPATH=$PWD:$PYTHONPATH
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw --format=csv > "$out/smi.csv" 2>&1
i=0
for cmd in "$@"; do
  i=$((i + 1))
  log=$(printf "%s/%02d.log" "$out" "$i")
  echo "== $cmd" > "$log"
  timeout "${BFLC_CMD_TIMEOUT:-600}" bash -c "$cmd" >> "$log" 2>&1
  rc=$?
  echo "== rc=$rc" >> "$log"
  echo "[$i] rc=$rc  $cmd"
  tail -n "${BFLC_TAIL:-6}" "$log" | cut -c1-400
done
