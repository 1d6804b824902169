"""ResNet-18 (CIFAR) convolution layers: implicit-GEMM path vs the im2col + GEMM path vs cuDNN
(bf16 channels_last), forward + backward, CUDA-event timed with an L2 flush between iterations.
    python scripts/conv_bench.py [--batch 64] [--iters 20]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as TF

from bflc_demo_b200.ops import nn as F

BF = torch.bfloat16
LAYERS = [  # cin, cout, k, stride, pad, hw
    (64, 64, 3, 1, 1, 32), (64, 128, 3, 2, 1, 32), (128, 128, 3, 1, 1, 16), (64, 128, 1, 2, 0, 32),
    (128, 256, 3, 2, 1, 16), (256, 256, 3, 1, 1, 8), (256, 512, 3, 2, 1, 8), (512, 512, 3, 1, 1, 4),
]


def timed(fn, iters, flush, graph=True):
    for _ in range(3):
        fn()
    if graph:   # replay a captured graph: device time of the kernels, not Python launch overhead
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            fn()
            with torch.cuda.graph(g, stream=st):
                fn()
        torch.cuda.synchronize()
        fn = g.replay
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return ts[len(ts) // 2] * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--layers", type=int, nargs="*", default=None, help="indices into LAYERS")
    ap.add_argument("--once", action="store_true", help="one eager implicit pass per layer (for ncu)")
    a = ap.parse_args()
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
    rows = []
    for li, (cin, cout, k, stride, pad, hw) in enumerate(LAYERS):
        if a.layers is not None and li not in a.layers:
            continue
        n = a.batch
        x = (torch.randn(n, hw, hw, cin, device="cuda") * 0.5).to(BF).requires_grad_(True)
        w = (torch.randn(cout, k * k * cin, device="cuda") * 0.05).to(BF)
        gw = torch.zeros(cout, k * k * cin, device="cuda")
        oh = (hw + 2 * pad - k) // stride + 1
        dy = torch.randn(n, oh, oh, cout, device="cuda").to(BF)

        def ours():
            x.grad = None
            y = F.conv2d(x, w, None, gw, None, k, k, stride, pad)
            y.backward(dy)

        xc = x.detach().permute(0, 3, 1, 2).requires_grad_(True)            # NCHW view, NHWC memory
        wc = w.view(cout, k, k, cin).permute(0, 3, 1, 2).detach().requires_grad_(True)
        dyc = dy.permute(0, 3, 1, 2)

        def cudnn():
            xc.grad = None; wc.grad = None
            y = TF.conv2d(xc, wc, None, stride=stride, padding=pad)
            y.backward(dyc)

        if a.once:
            ours(); ours()
            torch.cuda.synchronize()
            continue
        F._IMPLICIT = True
        t_imp = timed(ours, a.iters, flush)
        F._IMPLICIT = False
        t_exp = timed(ours, a.iters, flush)
        F._IMPLICIT = True
        t_dnn = timed(cudnn, a.iters, flush)
        flops = 3 * 2.0 * n * oh * oh * cout * k * k * cin
        rows.append(dict(layer=f"{cin}->{cout} k{k} s{stride} {hw}x{hw}", implicit_us=round(t_imp, 1),
                         im2col_us=round(t_exp, 1), cudnn_us=round(t_dnn, 1),
                         implicit_tflops=round(flops / t_imp / 1e6, 1),
                         speedup_vs_im2col=round(t_exp / t_imp, 2), vs_cudnn=round(t_dnn / t_imp, 2)))
        print(json.dumps(rows[-1]), flush=True)
    print("CONV_BENCH " + json.dumps(dict(batch=a.batch, rows=rows)))


if __name__ == "__main__":
    main()
