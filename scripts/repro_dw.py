"""Isolate the 'invalid argument' of the LeNet head weight-gradient GEMM (M=10 N=88 K=64)."""
import json, torch
from bflc_demo_b200.ops import gemm as G
from bflc_demo_b200._native import C
m = C()
out = {}
def attempt(name, fn):
    try:
        fn(); torch.cuda.synchronize(); out[name] = "ok"
    except Exception as e:  # noqa
        out[name] = str(e)[:160]
dl = torch.zeros(64, 16, device="cuda", dtype=torch.bfloat16)
h = torch.randn(64, 88, device="cuda").bfloat16()
gw = torch.zeros(10, 88, device="cuda")
attempt("encode_B_mn_N88_K64", lambda: m.gemm_b_map(h.data_ptr(), 88, 64, 88, True, False, 0, 64))
attempt("encode_B_mn_N96_K64", lambda: m.gemm_b_map(h.data_ptr(), 96, 64, 96, True, False, 0, 64))
attempt("dw_acc", lambda: G.gemm(dl[:, :10], h, out=gw, a_mn=True, b_mn=True, accumulate=True))
attempt("dw_plain", lambda: G.gemm(dl[:, :10], h, out=gw, a_mn=True, b_mn=True))
h96 = torch.randn(64, 96, device="cuda").bfloat16(); gw96 = torch.zeros(10, 96, device="cuda")
attempt("dw_N96", lambda: G.gemm(dl[:, :10], h96, out=gw96, a_mn=True, b_mn=True))
h128 = torch.randn(128, 88, device="cuda").bfloat16(); dl128 = torch.zeros(128, 16, device="cuda", dtype=torch.bfloat16)
attempt("dw_K128", lambda: G.gemm(dl128[:, :10], h128, out=gw, a_mn=True, b_mn=True))
dl64 = torch.zeros(64, 64, device="cuda", dtype=torch.bfloat16)
attempt("dw_lda64", lambda: G.gemm(dl64[:, :10], h, out=gw, a_mn=True, b_mn=True))
attempt("dw_M16", lambda: G.gemm(dl[:, :16], h, out=torch.zeros(16, 88, device="cuda"), a_mn=True, b_mn=True))
print("RESULT " + json.dumps(out))
