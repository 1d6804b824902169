"""Hot-cache, event-timed microbenchmarks of the individual kernels of one MLP training step
(and of a whole captured step), to separate launch overhead from execution time."""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import json, sys
import torch
from bflc_demo_b200.ops import gemm as G
from bflc_demo_b200._native import C
from bflc_demo_b200.models.mlp import FlatMLP, mlp_spec

def timeit(fn, n=200, warm=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    spec = mlp_spec(784, 256, 62)
    master = torch.empty(spec.total); spec.init_(master, 1); master = master.cuda()
    shadow = master.bfloat16(); grad = torch.zeros_like(master)
    tr = FlatMLP(spec, master, shadow, grad, B, lr=0.01)
    x = torch.rand(B, 784, device="cuda").bfloat16()
    y = torch.randint(0, 62, (B,), device="cuda", dtype=torch.int32)
    s, g, p = tr.s, tr.g, tr.p
    h, dl, dh = tr.h, tr.dlogits, tr.dh
    res = {}
    res["empty_launch_fill"] = timeit(lambda: C().fill_f32(grad[:8], 0.0))
    res["fwd1"] = timeit(lambda: G.gemm(x, s["w1"], out=h, bias=p["b1"], act=G.ACT_RELU))
    res["fwd1_nobias_f32out"] = timeit(lambda: G.gemm(x, s["w1"], out_dtype=torch.float32))
    res["xent"] = timeit(lambda: G.gemm_xent(h, s["w2"], y, n_classes=62, bias=p["b2"], dlogits=dl,
                         grad_scale=1.0 / B, loss_sum=tr.loss_sum, correct=tr.correct, colsum=g["b2"]))
    res["dW2_split%d" % tr.split_k] = timeit(lambda: G.gemm(dl[:, :62], h, out=g["w2"], a_mn=True, b_mn=True, split_k=tr.split_k))
    res["dW2_nosplit"] = timeit(lambda: G.gemm(dl[:, :62], h, out=g["w2"], a_mn=True, b_mn=True))
    res["dh"] = timeit(lambda: G.gemm(dl[:, :62], s["w2"], out=dh, b_mn=True, aux_in=h, act_bwd=1, colsum=g["b1"]))
    res["dh_plain"] = timeit(lambda: G.gemm(dl[:, :62], s["w2"], out=dh, b_mn=True))
    res["dW1_split%d" % tr.split_k] = timeit(lambda: G.gemm(dh, x, out=g["w1"], a_mn=True, b_mn=True, split_k=tr.split_k))
    res["dW1_nosplit"] = timeit(lambda: G.gemm(dh, x, out=g["w1"], a_mn=True, b_mn=True))
    res["optim"] = timeit(lambda: tr.optimizer_step(1))
    res["step_eager"] = timeit(lambda: (tr.forward_backward(x, y), tr.optimizer_step(1)), n=100)
    # one captured step
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        tr.forward_backward(x, y); tr.optimizer_step(1)
    st.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        for _ in range(8):
            tr.forward_backward(x, y); tr.optimizer_step(1)
    res["step_graph(8 steps)/8"] = timeit(lambda: gr.replay(), n=50) / 8
    # torch eager equivalents for context
    w1 = s["w1"]; 
    res["torch_fwd1"] = timeit(lambda: torch.relu(torch.addmm(p["b1"].bfloat16(), x, w1.t())))
    res["torch_dW1"] = timeit(lambda: dh.t() @ x)
    print("RESULT " + json.dumps({k: round(v, 2) for k, v in res.items()}))

main()
