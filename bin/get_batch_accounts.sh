#!/bin/bash
# Batch account generator (counterpart of the reference's python-sdk/bin/get_batch_accounts.sh):
#   bin/get_batch_accounts.sh <count> [out_dir]   ->  out_dir/node_<i>.pem (ECDSA secp256k1)
set -e
N=${1:?usage: get_batch_accounts.sh <count> [out_dir]}
OUT=${2:-accounts}
cd "$(dirname "$0")/.."
python - "$N" "$OUT" <<'PY'
import sys
from bflc_demo_b200.host.identity import generate_accounts
for i, a in generate_accounts(int(sys.argv[1]), sys.argv[2]).items():
    print(f"node_{i}.pem {a}")
PY
