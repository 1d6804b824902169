"""Small eager workloads for `ncu -k regex:<kernel>` captures (one GPU, no CUDA graph):
    python scripts/ncu_targets.py round      # 4 federated rounds, fp8 + Adam, eager launches
    python scripts/ncu_targets.py gemm2      # 8192^3 bf16 CTA-pair GEMM
    python scripts/ncu_targets.py mx8        # 8192^3 MXFP8 GEMM
    python scripts/ncu_targets.py attn       # BERT-base attention core fwd + bwd (batch 16, seq 128)
    python scripts/ncu_targets.py conv       # ResNet 128->128 3x3 16x16 batch 64 implicit-GEMM fwd + bwd
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

BF = torch.bfloat16


def main():
    what = sys.argv[1]
    if what == "round":
        from bflc_demo_b200.config import FLConfig
        from bflc_demo_b200.data.synthetic import femnist_like
        from bflc_demo_b200.engine.fused import FusedEngine
        cfg = FLConfig.for_world(1, hidden=256, batch_size=512, samples_per_client=4096, learning_rate=0.001,
                                 cuda_graph=False, dtype="fp8", optimizer="adam")
        eng = FusedEngine(cfg, femnist_like(1, 4096, seed=3)[0], rank=0, world=1, device=0)
        for _ in range(4):
            eng.run_round()
        torch.cuda.synchronize()
        assert not eng.drain_blocks()
    elif what == "gemm2":
        from bflc_demo_b200.ops import gemm as G
        a = (torch.randn(8192, 8192, device="cuda") * 0.1).to(BF)
        b = (torch.randn(8192, 8192, device="cuda") * 0.1).to(BF)
        for _ in range(3):
            G.gemm_2cta(a, b)
    elif what == "mx8":
        from bflc_demo_b200.ops.mx8 import gemm_mx8, quantize_mx8
        a = quantize_mx8((torch.randn(8192, 8192, device="cuda") * 0.1).to(BF))
        b = quantize_mx8((torch.randn(8192, 8192, device="cuda") * 0.1).to(BF))
        for _ in range(3):
            gemm_mx8(a, b)
    elif what == "attn":
        from bflc_demo_b200.ops import nn as F
        q, k, v = [(torch.randn(16 * 128, 768, device="cuda") * 0.5).to(BF).requires_grad_(True) for _ in range(3)]
        for _ in range(3):
            o = F.attention(q, k, v, 16, 128, 12)
            o.backward(torch.ones_like(o))
    elif what == "conv":
        from bflc_demo_b200.ops import nn as F
        x = (torch.randn(64, 16, 16, 128, device="cuda") * 0.5).to(BF).requires_grad_(True)
        w = (torch.randn(128, 9 * 128, device="cuda") * 0.05).to(BF)
        gw = torch.zeros(128, 9 * 128, device="cuda")
        for _ in range(3):
            y = F.conv2d(x, w, None, gw, None, 3, 3, 1, 1)
            y.backward(torch.ones_like(y))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
