"""Clock-stamp the phases of one GEMM CTA (bring-up instrumentation in gemm_sm100.cu):
cycles from kernel entry of CTA (0,0,0)."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bflc_demo_b200.ops import gemm as G
from bflc_demo_b200._native import C
m = C()
buf = torch.zeros(8, device="cuda", dtype=torch.int64)
def mk(*s): return (torch.randn(*s, device="cuda") * 0.5).bfloat16()
def plain(a, b, o, bn):
    M, K = a.shape; N = b.shape[0]
    return lambda: m.gemm(a, b, o, M, N, K, 1, K, K, 0, 0, False, False, False, 0, 1, N, 0, 1.0, None, 0, None, None,
                          0, None, 1, False, None, 0, 1.0, None, None, None, None, 0, 0, 0, 0, 0, bn)
a, b = mk(512, 784), mk(256, 784); bias = torch.randn(256, device="cuda")
o = torch.empty(512, 256, device="cuda", dtype=torch.bfloat16); of = torch.empty(512, 256, device="cuda")
a2, b2 = mk(4096, 2304), mk(256, 2304); o2 = torch.empty(4096, 256, device="cuda", dtype=torch.bfloat16)
a3, b3 = mk(1024, 4608), mk(512, 4608); o3 = torch.empty(1024, 512, device="cuda", dtype=torch.bfloat16)
cases = {"fwd1 512x256x784 bias relu": lambda: G.gemm(a, b, out=o, bias=bias, act=G.ACT_RELU),
         "plain 512x256x784 f32": lambda: G.gemm(a, b, out=of),
         "K=64 512x256x64": lambda: G.gemm(a[:, :64], b[:, :64], out=o),
         "4096x256x2304 bn64": plain(a2, b2, o2, 64), "4096x256x2304 bn128": plain(a2, b2, o2, 128),
         "4096x256x2304 bn256": plain(a2, b2, o2, 256),
         "1024x512x4608 bn64": plain(a3, b3, o3, 64), "1024x512x4608 bn256": plain(a3, b3, o3, 256)}
def conv_case(n, hw, cin, cout, mode):
    x = mk(n, hw, hw, cin); w = mk(cout, 9 * cin); dy = mk(n * hw * hw, cout)
    y = torch.empty(n * hw * hw, cout, device="cuda", dtype=torch.bfloat16)
    gw = torch.zeros(cout, 9 * cin, device="cuda"); dx = torch.empty(n * hw * hw, cin, device="cuda", dtype=torch.bfloat16)
    if mode == "fwd":
        return lambda: m.conv_gemm(1, 0, x, w, y, n, hw, hw, cin, hw, hw, 3, 3, 1, 1, cout, None, 0, None, None, 0, None, 1, False)
    if mode == "dgrad":
        return lambda: m.conv_gemm(1, 1, dy, w, dx, n, hw, hw, cout, hw, hw, 3, 3, 1, 1, cin, None, 0, None, None, 0, None, 1, False)
    sk = int(mode[5:])
    return lambda: m.conv_gemm(2, 0, x, dy, gw, n, hw, hw, cin, hw, hw, 3, 3, 1, 1, cout, None, 0, None, None, 0, None, sk, sk == 1)
for shp in ((64, 4, 512, 512), (64, 8, 256, 256), (64, 32, 64, 64)):
    for mode in ("fwd", "dgrad", "wgrad1", "wgrad4"):
        cases[f"conv {shp} {mode}"] = conv_case(*shp, mode)
res = {}
for name, fn in cases.items():
    for _ in range(5): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record(); fn(); ev[1].record(); torch.cuda.synchronize()
    m.set_debug_times(buf.data_ptr())
    fn(); torch.cuda.synchronize()
    m.set_debug_times(0)
    t = buf.cpu().tolist()
    res[name] = {"setup": t[1]-t[0], "first_tma_issued": t[2]-t[0], "first_full": t[3]-t[0],
                 "mma_all_issued": t[4]-t[0], "accum_ready": t[5]-t[0], "epilogue_done": t[6]-t[0],
                 "dealloc": t[7]-t[0], "kernel_us_warm_l2": round(ev[0].elapsed_time(ev[1]) * 1e3, 1)}
    print(name, json.dumps(res[name]), flush=True)
print("RESULT " + json.dumps(res))
