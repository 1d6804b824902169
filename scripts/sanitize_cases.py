"""One small federated case per process for compute-sanitizer (scripts/sanitize_gpu.sh): the
kernels with hand-rolled cross-proxy / cross-CTA synchronisation -- the persistent trainer
(bf16 and fp8, Adam, fused upload), the fp8 / bf16 validation chain, k_plan / k_consensus, the
input and blob quantisers -- at shapes small enough for racecheck.  Prints RESULT {json}."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bflc_demo_b200.config import FLConfig
from bflc_demo_b200.data.synthetic import femnist_like


def engine(dtype, optimizer, rounds=2):
    from bflc_demo_b200.engine.fused import FusedEngine
    cfg = FLConfig.for_world(1, model="mlp", hidden=256, batch_size=128, samples_per_client=256,
                             learning_rate=0.05 if optimizer == "sgd" else 1e-3, dtype=dtype,
                             optimizer=optimizer, cuda_graph=False)
    eng = FusedEngine(cfg, femnist_like(1, 256, seed=7, only=0)[0])
    for _ in range(rounds):
        eng.run_round()
    torch.cuda.synchronize()
    errs = eng.drain_blocks()
    st = eng.read_state()
    return dict(epoch=st["epoch"], loss=st["global_loss"], ledger_errs=errs, chain_ok=eng.host_ledger.verify_chain())


def main():
    case = sys.argv[1]
    dtype, opt = case.split("_")
    out = dict(case=case, **engine(dtype, opt))
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
