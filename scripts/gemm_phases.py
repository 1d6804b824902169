"""Clock-stamp the phases of one GEMM CTA (bring-up instrumentation in gemm_sm100.cu)."""
import json, sys, torch
from bflc_demo_b200.ops import gemm as G
from bflc_demo_b200._native import C
m = C()
buf = torch.zeros(8, device="cuda", dtype=torch.int64)
def mk(*s): return (torch.randn(*s, device="cuda") * 0.5).bfloat16()
cases = {"fwd1 512x256x784 bias relu": lambda: G.gemm(a, b, out=o, bias=bias, act=G.ACT_RELU),
         "plain 512x256x784 f32": lambda: G.gemm(a, b, out=of),
         "K=64 512x256x64": lambda: G.gemm(a[:, :64], b[:, :64], out=o)}
a, b = mk(512, 784), mk(256, 784); bias = torch.randn(256, device="cuda")
o = torch.empty(512, 256, device="cuda", dtype=torch.bfloat16); of = torch.empty(512, 256, device="cuda")
res = {}
for name, fn in cases.items():
    for _ in range(5): fn()
    torch.cuda.synchronize()
    m.set_debug_times(buf.data_ptr())
    fn(); torch.cuda.synchronize()
    m.set_debug_times(0)
    t = buf.cpu().tolist()
    res[name] = {"setup": t[1]-t[0], "first_tma_issued": t[2]-t[0], "first_full": t[3]-t[0],
                 "mma_all_issued": t[4]-t[0], "accum_ready": t[5]-t[0], "epilogue_done": t[6]-t[0],
                 "dealloc": t[7]-t[0]}
print("RESULT " + json.dumps(res))
