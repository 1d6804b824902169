"""Model families of BASELINE.json on the hand-written kernels: LeNet-5 (config #3),
ResNet-18 (config #4), BERT-base (config #5) and the generic MLP / softmax regression, all
as *functional* models over a flat parameter buffer (``models/flat.py``):

    net = LeNet5();  bound = net.bind(master, shadow, grad)
    loss = net.loss(bound, x, y);  loss.backward()        # grads land in the flat grad buffer
    hits = net.correct(bound_of_any_weights, x, y)        # e.g. a peer's uploaded weights

Channel / feature dims that would break TMA's 16-byte row alignment are padded to a multiple
of 8 with zero-initialised weights; a zero pad channel receives exactly zero gradient (its
activation is relu(0)=0 and the next layer's weights for it are zero), so the padded network
computes exactly the un-padded one.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import torch

from .._native import C
from ..ops import gemm as G
from ..ops import nn as F
from .flat import ParamSpec

BF = torch.bfloat16


def _up8(n: int) -> int:
    return (n + 7) // 8 * 8


@dataclass
class Bound:
    P: Dict[str, torch.Tensor]            # fp32 master views (biases, norm params, running stats)
    S: Dict[str, torch.Tensor]            # bf16 shadow views (GEMM operands)
    G: Optional[Dict[str, torch.Tensor]]  # fp32 grad views or None (inference)

    def g(self, name):
        return self.G[name] if self.G is not None else None


class FlatNet:
    """Base: owns a ParamSpec, binds flat buffers, provides loss/correct on top of ``logits_in``."""
    spec: ParamSpec
    n_classes: int

    def bind(self, master, shadow, grad=None) -> Bound:
        return Bound(self.spec.views(master), self.spec.views(shadow),
                     self.spec.views(grad) if grad is not None else None)

    def init_(self, master: torch.Tensor, seed: int = 0):
        self.spec.init_(master, seed)
        self._post_init(self.spec.views(master))

    def _post_init(self, P):  # zero the padding rows/cols, set norm scales
        pass

    def preprocess(self, x_raw: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def features(self, b: Bound, x, train: bool):  # -> [rows, feat] bf16 input of the head
        raise NotImplementedError

    head = ("fc.w", "fc.b")

    def loss(self, b: Bound, x, y, correct=None):
        h = self.features(b, x, True)
        w, bias = self.head
        return F.linear_xent(h, b.S[w], b.P[bias], b.g(w), b.g(bias), y, correct)

    @torch.no_grad()
    def correct(self, b: Bound, x, y) -> torch.Tensor:
        h = self.features(b, x, False)
        w, bias = self.head
        cnt = torch.zeros(1, device=h.device, dtype=torch.int32)
        G.gemm_argmax_acc(h.contiguous(), b.S[w], y, cnt, n_classes=self.n_classes, bias=b.P[bias])
        return cnt


# ----------------------------------------------------------------------------- MLP
class MLPNet(FlatNet):
    def __init__(self, in_dim=784, hidden=256, n_classes=62):
        self.n_classes = n_classes
        self.in_dim = in_dim
        self.spec = ParamSpec([("fc1.w", (hidden, in_dim)), ("fc1.b", (hidden,)),
                               ("fc.w", (n_classes, hidden)), ("fc.b", (n_classes,))])

    def preprocess(self, x_raw):
        x = x_raw.reshape(x_raw.shape[0], -1).contiguous()
        out = torch.empty(x.shape, device=x.device, dtype=BF)
        C().cast_u8_to_bf16(x, out, 1.0 / 255.0)
        return out

    def features(self, b, x, train):
        if train:
            x = x.detach().requires_grad_(True)
        return F.linear(x, b.S["fc1.w"], b.P["fc1.b"], b.g("fc1.w"), b.g("fc1.b"), G.ACT_RELU,
                        need_dx=False)


# ----------------------------------------------------------------------------- LeNet-5
class LeNet5(FlatNet):
    """conv5x5(3->6) - pool - conv5x5(6->16) - pool - fc120 - fc84 - fc10 on 3x32x32 inputs.
    6 -> 8 channels and 84 -> 88 features are zero-padded (see module docstring)."""

    def __init__(self, n_classes=10, in_ch=3):
        self.n_classes, self.in_ch = n_classes, in_ch
        self.c1, self.c2 = 8, 16          # 6 (+2 pad), 16
        self.k1 = _up8(25 * in_ch)        # 75 -> 80
        self.k2 = 25 * self.c1            # 200
        self.f1, self.f2 = 120, 88        # 84 (+4 pad)
        self.flat = 5 * 5 * self.c2       # 400
        self.spec = ParamSpec([
            ("conv1.w", (self.c1, self.k1)), ("conv1.b", (self.c1,)),
            ("conv2.w", (self.c2, self.k2)), ("conv2.b", (self.c2,)),
            ("fc1.w", (self.f1, self.flat)), ("fc1.b", (self.f1,)),
            ("fc2.w", (self.f2, self.f1)), ("fc2.b", (self.f2,)),
            ("fc.w", (n_classes, self.f2)), ("fc.b", (n_classes,))])

    def _post_init(self, P):
        P["conv1.w"][6:].zero_(); P["conv1.w"][:, 25 * self.in_ch:].zero_()
        w2 = P["conv2.w"].view(self.c2, 25, self.c1)
        w2[:, :, 6:].zero_()
        P["fc2.w"][84:].zero_(); P["fc.w"][:, 84:].zero_()

    def preprocess(self, x_raw):  # uint8 [N, 3, 32, 32] -> bf16 NHWC
        x = x_raw.permute(0, 2, 3, 1).contiguous()
        out = torch.empty(x.shape, device=x.device, dtype=BF)
        C().cast_u8_to_bf16(x.view(-1), out.view(-1), 1.0 / 255.0)
        return out

    def features(self, b, x, train):
        if train:
            x = x.detach().requires_grad_(True)
        g = b.g
        x = F.conv2d(x, b.S["conv1.w"], b.P["conv1.b"], g("conv1.w"), g("conv1.b"), 5, 5, 1, 0,
                     G.ACT_RELU, need_dx=False)
        x = F.maxpool2d(x, 2, 2)
        x = F.conv2d(x, b.S["conv2.w"], b.P["conv2.b"], g("conv2.w"), g("conv2.b"), 5, 5, 1, 0,
                     G.ACT_RELU)
        x = F.maxpool2d(x, 2, 2)
        x = x.reshape(x.shape[0], -1)
        x = F.linear(x, b.S["fc1.w"], b.P["fc1.b"], g("fc1.w"), g("fc1.b"), G.ACT_RELU)
        return F.linear(x, b.S["fc2.w"], b.P["fc2.b"], g("fc2.w"), g("fc2.b"), G.ACT_RELU)


# ----------------------------------------------------------------------------- ResNet-18
class ResNet18(FlatNet):
    """CIFAR-style ResNet-18: conv3x3(3->64) stem, stages [64,128,256,512] x 2 BasicBlocks,
    global average pool, fc.  ~11.2 M parameters.  Batch-norm running statistics live in the
    flat buffer (so FedAvg averages them like every other parameter)."""

    def __init__(self, n_classes=10, in_ch=3, widths=(64, 128, 256, 512)):
        self.n_classes, self.in_ch, self.widths = n_classes, in_ch, widths
        ents: List[Tuple[str, Tuple[int, ...]]] = []

        def bn(name, c):
            ents.extend([(f"{name}.gamma", (c,)), (f"{name}.beta", (c,)),
                         (f"{name}.rmean", (c,)), (f"{name}.rvar", (c,))])

        self.k_stem = _up8(9 * in_ch)
        ents.append(("stem.w", (widths[0], self.k_stem)))
        bn("stem.bn", widths[0])
        self.blocks = []
        cin = widths[0]
        for si, c in enumerate(widths):
            for bi in range(2):
                stride = 2 if (si > 0 and bi == 0) else 1
                name = f"l{si}.{bi}"
                ents.append((f"{name}.c1.w", (c, 9 * cin)))
                bn(f"{name}.bn1", c)
                ents.append((f"{name}.c2.w", (c, 9 * c)))
                bn(f"{name}.bn2", c)
                down = stride != 1 or cin != c
                if down:
                    ents.append((f"{name}.down.w", (c, cin)))
                    bn(f"{name}.dbn", c)
                self.blocks.append((name, cin, c, stride, down))
                cin = c
        ents.extend([("fc.w", (n_classes, widths[-1])), ("fc.b", (n_classes,))])
        self.spec = ParamSpec(ents)

    def _post_init(self, P):
        for k, v in P.items():
            if k.endswith(".rvar"):
                v.fill_(1.0)
        P["stem.w"][:, 9 * self.in_ch:].zero_()

    preprocess = LeNet5.preprocess

    def _bn(self, b, name, x, train, relu, residual=None):
        return F.batchnorm(x, b.P[f"{name}.gamma"], b.P[f"{name}.beta"], b.g(f"{name}.gamma"),
                           b.g(f"{name}.beta"), b.P[f"{name}.rmean"], b.P[f"{name}.rvar"],
                           training=train, relu=relu, residual=residual)

    def features(self, b, x, train):
        if train:
            x = x.detach().requires_grad_(True)
        g = b.g
        x = F.conv2d(x, b.S["stem.w"], None, g("stem.w"), None, 3, 3, 1, 1, need_dx=False)
        x = self._bn(b, "stem.bn", x, train, True)
        for name, cin, c, stride, down in self.blocks:
            idt = x
            y = F.conv2d(x, b.S[f"{name}.c1.w"], None, g(f"{name}.c1.w"), None, 3, 3, stride, 1)
            y = self._bn(b, f"{name}.bn1", y, train, True)
            y = F.conv2d(y, b.S[f"{name}.c2.w"], None, g(f"{name}.c2.w"), None, 3, 3, 1, 1)
            if down:
                idt = F.conv2d(x, b.S[f"{name}.down.w"], None, g(f"{name}.down.w"), None, 1, 1,
                               stride, 0)
                idt = self._bn(b, f"{name}.dbn", idt, train, False)
            x = self._bn(b, f"{name}.bn2", y, train, True, residual=idt)
        return F.global_avgpool(x)


# ----------------------------------------------------------------------------- BERT-base
class BertBase(FlatNet):
    """BERT-base encoder for sequence classification: 12 layers, hidden 768, 12 heads, FFN 3072,
    vocab 30522, 512 positions (seq_len 128 in config #5); post-LN, GELU; classifier on [CLS]
    through a 768->768 GELU pooler.  ~109 M parameters."""
    head = ("cls.w", "cls.b")

    def __init__(self, n_classes=2, layers=12, hidden=768, heads=12, ffn=3072, vocab=30522,
                 max_pos=512):
        self.n_classes, self.L, self.Hd, self.heads, self.ffn = n_classes, layers, hidden, heads, ffn
        ents: List[Tuple[str, Tuple[int, ...]]] = [
            ("emb.word", (vocab, hidden)), ("emb.pos", (max_pos, hidden)),
            ("emb.ln.gamma", (hidden,)), ("emb.ln.beta", (hidden,))]
        for i in range(layers):
            p = f"enc{i}"
            for nm in ("q", "k", "v", "o"):
                ents.extend([(f"{p}.{nm}.w", (hidden, hidden)), (f"{p}.{nm}.b", (hidden,))])
            ents.extend([(f"{p}.ln1.gamma", (hidden,)), (f"{p}.ln1.beta", (hidden,)),
                         (f"{p}.ff1.w", (ffn, hidden)), (f"{p}.ff1.b", (ffn,)),
                         (f"{p}.ff2.w", (hidden, ffn)), (f"{p}.ff2.b", (hidden,)),
                         (f"{p}.ln2.gamma", (hidden,)), (f"{p}.ln2.beta", (hidden,))])
        ents.extend([("pool.w", (hidden, hidden)), ("pool.b", (hidden,)),
                     ("cls.w", (n_classes, hidden)), ("cls.b", (n_classes,))])
        self.spec = ParamSpec(ents)

    def _post_init(self, P):
        P["emb.word"].mul_(0.02 * (self.Hd ** 0.5))  # ~N(0, 0.02)-scale embeddings
        P["emb.pos"].mul_(0.02 * (self.Hd ** 0.5))

    def preprocess(self, x_raw):  # int64 [N, S] -> int32
        return x_raw.to(torch.int32).contiguous()

    def _lin(self, b, name, x, act=G.ACT_NONE):
        return F.linear(x, b.S[f"{name}.w"], b.P[f"{name}.b"], b.g(f"{name}.w"), b.g(f"{name}.b"), act)

    def _ln(self, b, name, x):
        return F.layernorm(x, b.P[f"{name}.gamma"], b.P[f"{name}.beta"], b.g(f"{name}.gamma"),
                           b.g(f"{name}.beta"))

    def features(self, b, ids, train):
        B, S = ids.shape
        x = F.embedding(ids.reshape(-1), b.S["emb.word"], b.S["emb.pos"], b.g("emb.word"),
                        b.g("emb.pos"), S)
        x = self._ln(b, "emb.ln", x)
        for i in range(self.L):
            p = f"enc{i}"
            q, k, v = (self._lin(b, f"{p}.{nm}", x) for nm in ("q", "k", "v"))
            a = F.attention(q, k, v, B, S, self.heads)
            x = self._ln(b, f"{p}.ln1", F.add(x, self._lin(b, f"{p}.o", a)))
            h = self._lin(b, f"{p}.ff1", x, G.ACT_GELU)
            x = self._ln(b, f"{p}.ln2", F.add(x, self._lin(b, f"{p}.ff2", h)))
        cls_tok = x.view(B, S, self.Hd)[:, 0, :]
        return self._lin(b, "pool", cls_tok, G.ACT_GELU)


def build_model(name: str, n_classes: int, **kw) -> FlatNet:
    name = name.lower()
    if name == "mlp":
        return MLPNet(kw.get("in_dim", 784), kw.get("hidden", 256), n_classes)
    if name in ("lenet5", "lenet"):
        return LeNet5(n_classes)
    if name == "resnet18":
        return ResNet18(n_classes)
    if name in ("bert", "bert-base", "bert_base"):
        return BertBase(n_classes, layers=kw.get("layers", 12))
    raise ValueError(f"unknown model {name}")
