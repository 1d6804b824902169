"""Where does the implicit-GEMM forward lose time?  Same GEMM shape through (a) the 4-D box
producer, (b) the plain 2-D producer on a materialised im2col matrix, per N-tile width."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bflc_demo_b200._native import C
BF = torch.bfloat16


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        flush.zero_(); a.record(); fn(); b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    return round(ts[len(ts) // 2] * 1e3, 1)


for (n, hw, cin, cout) in [(64, 8, 256, 256), (64, 32, 64, 64), (64, 4, 512, 512), (16, 32, 256, 256)]:
    x = (torch.randn(n, hw, hw, cin, device="cuda") * 0.5).to(BF)
    w = (torch.randn(cout, 9 * cin, device="cuda") * 0.05).to(BF)
    rows = n * hw * hw
    y = torch.empty(rows, cout, device="cuda", dtype=BF)
    col = torch.empty(rows, 9 * cin, device="cuda", dtype=BF)
    C().im2col(x, col, n, cin, hw, hw, 3, 3, 1, 1, hw, hw)
    out = dict(shape=f"n{n} {hw}x{hw} {cin}->{cout}", M=rows, N=cout, K=9 * cin)
    out["implicit"] = timed(lambda: C().conv_gemm(1, 0, x, w, y, n, hw, hw, cin, hw, hw, 3, 3, 1, 1, cout, None, 0,
                                                   None, None, 0, None, 1, False))
    for bn in (64, 128, 256):
        if bn > cout * 2:
            continue
        out[f"plain_bn{bn}"] = timed(lambda: C().gemm(col, w, y, rows, cout, 9 * cin, 1, 9 * cin, 9 * cin, 0, 0,
                                                       False, False, False, 0, 1, cout, 0, 1.0, None, 0, None, None,
                                                       0, None, 1, False, None, 0, 1.0, None, None, None, None,
                                                       0, 0, 0, 0, 0, bn))
    yc = torch.empty_like(y)
    out["cublas"] = timed(lambda: torch.mm(col, w.t(), out=yc))
    print(json.dumps(out), flush=True)
