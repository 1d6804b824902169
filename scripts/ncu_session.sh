#!/bin/bash
# Round-2 ncu evidence on ONE GPU: a launch list of a federated round and one `--set full` capture
# per hot kernel (clock control off, source import on).  Reports land in gpurun_out/<tag>/.
TAG=${1:-ncu_r2}
OUT=gpurun_out/$TAG
mkdir -p "$OUT"
export PYTHONPATH=$PWD:$PYTHONPATH BFLC_NO_AUTOBUILD=1
NCU="ncu --clock-control none"
cap() {  # name, kernel regex, skip, target, [extra ncu flags]
  timeout 400 $NCU --set full $5 -k "regex:$2" --launch-skip "$3" --launch-count 1 \
      -f -o "$OUT/$1" python scripts/ncu_targets.py "$4" > "$OUT/$1.log" 2>&1
  echo "$1 rc=$?"
}
timeout 300 $NCU --metrics gpu__time_duration.sum --csv --log-file "$OUT/launches_round.csv" \
    python scripts/ncu_targets.py round > "$OUT/launches_round.log" 2>&1
echo "launch list rc=$?"
cap mlp_round   "mlp_round_kernel"  2 round "--import-source on"
cap mlp_val     "mlp_val_kernel"    2 round
cap consensus   "k_consensus"       2 round "--import-source on"
cap prep_inputs "k_prep_inputs"     2 round
cap gemm2       "gemm2_kernel"      1 gemm2
cap gemm_mx8    "gemm_mx8_kernel"   1 mx8
cap attn_fwd    "attn_fwd_kernel"   1 attn
cap attn_bwd    "attn_bwd_kernel"   1 attn
cap conv_fwd    "gemm_kernel"       3 conv
du -sh "$OUT"; ls -la "$OUT"
