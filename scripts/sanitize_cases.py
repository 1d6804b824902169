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


def conv_attn(case):
    """The round-2 kernels outside the flagship round: implicit-GEMM convolution (stride 1 and 2,
    forward + both gradients) and fused attention forward / backward, at small shapes."""
    from bflc_demo_b200.ops import nn as F
    BF = torch.bfloat16
    torch.manual_seed(0)
    if case == "conv":
        tot = 0.0
        for (n, hw, cin, cout, stride) in ((2, 16, 64, 64, 1), (2, 16, 64, 128, 2), (3, 4, 128, 64, 1)):
            x = (torch.randn(n, hw, hw, cin, device="cuda") * 0.5).to(BF).requires_grad_(True)
            w = (torch.randn(cout, 9 * cin, device="cuda") * 0.05).to(BF)
            gw = torch.zeros(cout, 9 * cin, device="cuda")
            y = F.conv2d(x, w, None, gw, None, 3, 3, stride, 1)
            y.backward(torch.ones_like(y))
            tot += float(gw.abs().sum()) + float(x.grad.float().abs().sum())
        return dict(checksum=tot)
    q, k, v = [(torch.randn(2 * 128, 2 * 64, device="cuda") * 0.5).to(BF).requires_grad_(True) for _ in range(3)]
    o = F.attention(q, k, v, 2, 128, 2)
    o.backward(torch.ones_like(o))
    return dict(checksum=float(o.float().abs().sum()) + float(q.grad.float().abs().sum()))


def main():
    case = sys.argv[1]
    if case in ("conv", "attn"):
        out = dict(case=case, **conv_attn(case))
        torch.cuda.synchronize()
        print("RESULT " + json.dumps(out))
        return
    dtype, opt = case.split("_")
    out = dict(case=case, **engine(dtype, opt))
    print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
