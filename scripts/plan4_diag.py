"""Diagnostic for the experimental cluster plan (BFLC_MLP_EXPERIMENTAL=1): per-tensor differences
between plan 3 and plan 4 after one and after three steps."""
import json, os
import torch
from bflc_demo_b200.models.mlp import FlatMLP, mlp_spec

def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()

def run(plan, B, steps):
    spec = mlp_spec(784, 256, 62)
    init = torch.empty(spec.total); spec.init_(init, seed=2)
    torch.manual_seed(11)
    X = torch.rand(B * steps, 784, device="cuda").bfloat16()
    Y = torch.randint(0, 62, (B * steps,), device="cuda", dtype=torch.int32)
    master = init.cuda().clone(); shadow = master.bfloat16(); grad = torch.zeros_like(master)
    tr = FlatMLP(spec, master, shadow, grad, B, lr=0.05)
    bar = torch.zeros(1, device="cuda", dtype=torch.int32)
    tr.train_epoch_fused(X, Y, steps, bar.data_ptr(), None, plan, 1)
    torch.cuda.synchronize()
    v = spec.views(master)
    return {k: v[k].clone() for k in ("w1", "b1", "w2", "b2")}, tr.loss_sum.item(), int(tr.correct.item()), tr.h.clone(), tr.dh.clone(), tr.dlogits.clone()

out = {}
for B, steps in ((256, 1), (256, 3), (512, 2)):
    a = run(3, B, steps); b = run(4, B, steps)
    out[f"B{B}_s{steps}"] = {**{k: rel(b[0][k], a[0][k]) for k in a[0]}, "loss": [a[1], b[1]], "correct": [a[2], b[2]],
                             "h": rel(b[3], a[3]), "dh": rel(b[4], a[4]), "dlogits": rel(b[5], a[5])}
print("PLAN4 " + json.dumps(out))
