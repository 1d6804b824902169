"""Autograd layer functions over the hand-written kernels.

Activations are bf16 autograd tensors; parameters are NOT autograd leaves: every function takes
the bf16 shadow weight it computes with, the fp32 master bias, and the fp32 gradient views
(``gw``/``gb``, slices of the model's flat gradient buffer) that its backward accumulates into
directly.  Passing ``gw=None`` gives an inference-only call, and because a weight is just a
pointer, the same functions run a model whose parameters live in a *peer GPU's* HBM (committee
validation: the GEMMs' TMA loads pull the candidate's weights over NVLink).

Every GEMM here is ``ops.gemm`` (tcgen05).  Reference ops covered: K1 matmul+bias, K2
softmax-xent, K3 backward; the rest exists for the LeNet/ResNet/BERT configs.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch.autograd import Function

from .._native import C
from . import gemm as G

BF = torch.bfloat16

# Forward-GEMM precision of Linear / Conv2d: "bf16" (kind::f16) or "mx8" (block-scaled fp8,
# kind::mxf8f6f4.block_scale: activations and weights are quantised to e4m3 with one UE8M0
# scale per 32 K-elements right before the GEMM).  Backward GEMMs stay bf16 on the saved bf16
# operands (forward-fp8 / backward-bf16 recipe); master weights, gradients and optimizer fp32.
_PRECISION = "bf16"
_MX_BUF: dict = {}


def set_precision(p: str) -> str:
    """-> the previous setting."""
    global _PRECISION
    if p not in ("bf16", "mx8"):
        raise ValueError("precision must be 'bf16' or 'mx8'")
    prev, _PRECISION = _PRECISION, p
    return prev


def get_precision() -> str:
    return _PRECISION


def _mx_quant(tag: str, t: torch.Tensor):
    from .mx8 import MX8, quantize_mx8
    R, K = t.shape
    key = (tag, R, K, t.device.index)
    buf = _MX_BUF.get(key)
    if buf is None:
        ld = (K + 15) // 16 * 16
        buf = MX8(torch.zeros(R, ld, device=t.device, dtype=torch.float8_e4m3fn),
                  torch.empty(C().mx8_sf_bytes(R, K), device=t.device, dtype=torch.uint8), R, K)
        _MX_BUF[key] = buf
    return quantize_mx8(t, out=buf)


def _fwd_gemm(x, w, y, bias, act, pre):
    """y = act(x @ w^T + bias) in the configured forward precision."""
    if _PRECISION == "mx8" and act != G.ACT_GELU and y.stride(0) % 4 == 0:
        from .mx8 import gemm_mx8
        gemm_mx8(_mx_quant("x", x), _mx_quant("w", w), out=y, bias=bias, act=act)
    else:
        G.gemm(x, w, out=y, bias=bias, act=act, aux_out=pre)


def _split_k(out_rows: int, out_cols: int, k: int) -> int:
    """Split the reduction when a weight-gradient GEMM has few output tiles but a long K."""
    tiles = ((out_rows + 127) // 128) * ((out_cols + 255) // 256)
    kb = (k + 63) // 64
    if tiles >= 74 or kb < 8:
        return 1
    return max(1, min(kb // 2, 148 // tiles, 32))


def _dw(dz: torch.Tensor, x: torch.Tensor, gw: torch.Tensor):
    """gw[N, K] += dz[M, N]^T @ x[M, K]   (both operands consumed MN-major, no transposes)."""
    sk = _split_k(gw.shape[0], gw.shape[1], dz.shape[0])
    if sk > 1:
        G.gemm(dz, x, out=gw, a_mn=True, b_mn=True, split_k=sk)
    else:
        G.gemm(dz, x, out=gw, a_mn=True, b_mn=True, accumulate=True)


class LinearFn(Function):
    @staticmethod
    def forward(ctx, x, w, b, gw, gb, act, need_dx=True):
        ctx.need_dx = need_dx
        x = x.contiguous()
        M, N = x.shape[0], w.shape[0]
        y = torch.empty(M, N, device=x.device, dtype=BF)
        pre = torch.empty_like(y) if act == G.ACT_GELU else None
        _fwd_gemm(x, w, y, b, act, pre)
        ctx.save_for_backward(x, w, y if act == G.ACT_RELU else pre)
        ctx.gw, ctx.gb, ctx.act = gw, gb, act
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, aux = ctx.saved_tensors
        dy = dy.contiguous()
        M, N = dy.shape
        if ctx.act != G.ACT_NONE:
            dz = torch.empty_like(dy)
            C().act_bwd_colsum(dy, aux, dz, ctx.gb, M, N, ctx.act)
        else:
            dz = dy
            if ctx.gb is not None:
                C().act_bwd_colsum(dy, None, None, ctx.gb, M, N, 0)
        if ctx.gw is not None:
            _dw(dz, x, ctx.gw)
        dx = G.gemm(dz, w, b_mn=True) if (ctx.needs_input_grad[0] and ctx.need_dx) else None
        return dx, None, None, None, None, None, None


def linear(x, w, b=None, gw=None, gb=None, act=G.ACT_NONE, need_dx=True):
    """``need_dx=False`` on a model's first layer: its input only carries ``requires_grad`` so
    that autograd runs the backward functions (parameters are not autograd leaves)."""
    return LinearFn.apply(x, w, b, gw, gb, act, need_dx)


class LinearXentFn(Function):
    """Classifier head fused with softmax-cross-entropy (mean over rows); also counts hits."""

    @staticmethod
    def forward(ctx, h, w, b, gw, gb, labels, correct):
        h = h.contiguous()
        M, n_cls = h.shape[0], w.shape[0]
        ncp = (n_cls + 7) // 8 * 8
        dl = torch.zeros(M, ncp, device=h.device, dtype=BF)
        loss = torch.zeros(1, device=h.device, dtype=torch.float32)
        G.gemm_xent(h, w, labels, n_classes=n_cls, bias=b, dlogits=dl, grad_scale=1.0 / M,
                    loss_sum=loss, correct=correct, colsum=gb)
        ctx.save_for_backward(h, w, dl)
        ctx.gw, ctx.n_cls = gw, n_cls
        return loss / M

    @staticmethod
    def backward(ctx, gout):
        h, w, dl = ctx.saved_tensors
        dlv = dl[:, :ctx.n_cls]
        if ctx.gw is not None:
            _dw(dlv, h, ctx.gw)
        dh = G.gemm(dlv, w, b_mn=True) if ctx.needs_input_grad[0] else None
        return dh, None, None, None, None, None, None


def linear_xent(h, w, b, gw, gb, labels, correct=None):
    return LinearXentFn.apply(h, w, b, gw, gb, labels, correct)


_IMPLICIT = os.environ.get("BFLC_CONV_IMPLICIT", "1") != "0"


def _pix_tile(pix: int, OH: int, OW: int) -> bool:
    """Can `pix` consecutive output pixels be fetched as one TMA box of whole image rows?"""
    if OW > pix or pix % OW:
        return False
    rows = pix // OW
    return OH % rows == 0 if rows <= OH else rows % OH == 0


def conv_is_implicit(H, W, Cin, kh, kw, stride, pad, Kp) -> bool:
    """Implicit-GEMM eligibility (csrc/kernels/gemm_sm100.cu, ConvView): 64-channel K blocks and
    pixel tiles made of whole image rows; everything else takes the im2col path."""
    OH, OW = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
    return (_IMPLICIT and _PRECISION == "bf16" and Cin % 64 == 0 and Kp == kh * kw * Cin
            and _pix_tile(128, OH, OW) and _pix_tile(64, OH, OW) and OW * stride <= 256)


def _conv_rows_gemm(flip, act_src, w, out, N, GH, GW, Cc, OH, OW, kh, kw, stride, pad, n_out, bias=None,
                    act=G.ACT_NONE, pre=None):
    """Mode-1 implicit GEMM (rows = pixels).  Few output tiles but a long reduction (the deep,
    small-image ResNet layers) would leave most SMs idle and run the rest at the 128 x 64 tile's
    L2-feed limit: split the taps x channels reduction over CTAs into an fp32 workspace
    (red.add), then cast -- wide 256-column tiles on every SM."""
    M, kb = N * OH * OW, kh * kw * Cc // 64
    tiles = ((M + 127) // 128) * ((n_out + 255) // 256)
    sk = min(148 // tiles, kb // 4) if tiles <= 74 else 1
    if sk >= 2 and bias is None and act == G.ACT_NONE and pre is None and n_out % 4 == 0:
        ws = torch.zeros(M, n_out, device=out.device, dtype=torch.float32)
        C().conv_gemm(1, flip, act_src, w, ws, N, GH, GW, Cc, OH, OW, kh, kw, stride, pad, n_out, None, 0,
                      None, None, 0, None, sk, False)
        C().cast_f32_to_bf16(ws.view(-1), out.view(-1))
    else:
        C().conv_gemm(1, flip, act_src, w, out, N, GH, GW, Cc, OH, OW, kh, kw, stride, pad, n_out, bias, act,
                      pre, None, 0, None, 1, False)


class ConvImplicitFn(Function):
    """NHWC convolution as an implicit GEMM: the tcgen05 GEMM's TMA producer fetches each
    (filter tap, 64 channels) K block as a tap-shifted 4-D box of the activation itself, the
    border zero-filled by the TMA unit -- no im2col buffer in forward, input-gradient
    (stride 1) or weight-gradient."""

    @staticmethod
    def forward(ctx, x, w, b, gw, gb, kh, kw, stride, pad, act, need_dx=True):
        x = x.contiguous()
        N, H, W, Cin = x.shape
        Cout = w.shape[0]
        OH, OW = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
        y = torch.empty(N * OH * OW, Cout, device=x.device, dtype=BF)
        pre = torch.empty_like(y) if act == G.ACT_GELU else None
        _conv_rows_gemm(0, x, w, y, N, H, W, Cin, OH, OW, kh, kw, stride, pad, Cout, b, act, pre)
        ctx.save_for_backward(x, w, y if act == G.ACT_RELU else pre)
        ctx.gw, ctx.gb, ctx.act, ctx.need_dx = gw, gb, act, need_dx
        ctx.geom = (N, Cin, H, W, kh, kw, stride, pad, OH, OW)
        return y.view(N, OH, OW, Cout)

    @staticmethod
    def backward(ctx, dy):
        x, w, aux = ctx.saved_tensors
        N, Cin, H, W, kh, kw, stride, pad, OH, OW = ctx.geom
        Cout = w.shape[0]
        dy = dy.contiguous().view(-1, Cout)
        rows = dy.shape[0]
        if ctx.act != G.ACT_NONE:
            dz = torch.empty_like(dy)
            C().act_bwd_colsum(dy, aux, dz, ctx.gb, rows, Cout, ctx.act)
        else:
            dz = dy
            if ctx.gb is not None:
                C().act_bwd_colsum(dy, None, None, ctx.gb, rows, Cout, 0)
        if ctx.gw is not None:
            tiles = ((Cout + 127) // 128) * ((kh * kw * Cin + 127) // 128)
            sk = max(1, min(148 // tiles, (rows // 64) // 2, 32))
            C().conv_gemm(2, 0, x, dz, ctx.gw, N, H, W, Cin, OH, OW, kh, kw, stride, pad, Cout, None, 0,
                          None, None, 0, None, sk, sk == 1)
        dx = None
        if ctx.needs_input_grad[0] and ctx.need_dx:
            dx = torch.empty(N, H, W, Cin, device=dy.device, dtype=BF)
            if Cout % 64 == 0 and _pix_tile(128, H, W):
                if stride == 1:
                    g, GH, GW = dz, OH, OW
                else:   # zero-stuffed dy on the input grid, then the stride-1 form
                    g, GH, GW = torch.empty(N, H, W, Cout, device=dy.device, dtype=BF), H, W
                    C().upsample_zero(dz, g, N, H, W, OH, OW, Cout, stride)
                _conv_rows_gemm(1, g, w, dx.view(-1, Cin), N, GH, GW, Cout, H, W, kh, kw, 1, pad, Cin)
            else:
                dcol = G.gemm(dz, w, b_mn=True)
                C().col2im(dcol, dx, N, Cin, H, W, kh, kw, stride, pad, OH, OW)
        return (dx,) + (None,) * 10


class Conv2dFn(Function):
    """NHWC convolution = im2col + tcgen05 GEMM (+bias, +activation epilogue)."""

    @staticmethod
    def forward(ctx, x, w, b, gw, gb, kh, kw, stride, pad, act, need_dx=True):
        ctx.need_dx = need_dx
        x = x.contiguous()
        N, H, W, Cin = x.shape
        Cout, Kp = w.shape
        OH, OW = (H + 2 * pad - kh) // stride + 1, (W + 2 * pad - kw) // stride + 1
        rows, kc = N * OH * OW, kh * kw * Cin
        col = (torch.zeros if Kp != kc else torch.empty)(rows, Kp, device=x.device, dtype=BF)
        C().im2col(x, col, N, Cin, H, W, kh, kw, stride, pad, OH, OW)
        y = torch.empty(rows, Cout, device=x.device, dtype=BF)
        pre = torch.empty_like(y) if act == G.ACT_GELU else None
        _fwd_gemm(col, w, y, b, act, pre)
        ctx.save_for_backward(col, w, y if act == G.ACT_RELU else pre)
        ctx.gw, ctx.gb, ctx.act = gw, gb, act
        ctx.geom = (N, Cin, H, W, kh, kw, stride, pad, OH, OW)
        return y.view(N, OH, OW, Cout)

    @staticmethod
    def backward(ctx, dy):
        col, w, aux = ctx.saved_tensors
        N, Cin, H, W, kh, kw, stride, pad, OH, OW = ctx.geom
        Cout = w.shape[0]
        dy = dy.contiguous().view(-1, Cout)
        rows = dy.shape[0]
        if ctx.act != G.ACT_NONE:
            dz = torch.empty_like(dy)
            C().act_bwd_colsum(dy, aux, dz, ctx.gb, rows, Cout, ctx.act)
        else:
            dz = dy
            if ctx.gb is not None:
                C().act_bwd_colsum(dy, None, None, ctx.gb, rows, Cout, 0)
        if ctx.gw is not None:
            _dw(dz, col, ctx.gw)
        dx = None
        if ctx.needs_input_grad[0] and ctx.need_dx:
            dcol = G.gemm(dz, w, b_mn=True)
            dx = torch.empty(N, H, W, Cin, device=dy.device, dtype=BF)
            C().col2im(dcol, dx, N, Cin, H, W, kh, kw, stride, pad, OH, OW)
        return (dx,) + (None,) * 10


def conv2d(x, w, b, gw, gb, kh, kw, stride=1, pad=0, act=G.ACT_NONE, need_dx=True):
    if conv_is_implicit(x.shape[1], x.shape[2], x.shape[3], kh, kw, stride, pad, w.shape[1]):
        return ConvImplicitFn.apply(x, w, b, gw, gb, kh, kw, stride, pad, act, need_dx)
    return Conv2dFn.apply(x, w, b, gw, gb, kh, kw, stride, pad, act, need_dx)


class BatchNormFn(Function):
    """Channels-last batch norm over [rows, C] with fused (+residual) (+ReLU)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, ggamma, gbeta, run_mean, run_var, training, relu, residual):
        shape = x.shape
        Cc = shape[-1]
        x2 = x.contiguous().view(-1, Cc)
        rows = x2.shape[0]
        y = torch.empty_like(x2)
        if training:
            mean = torch.empty(Cc, device=x.device, dtype=torch.float32)
            rstd = torch.empty_like(mean)
        else:
            mean = run_mean.clone()
            rstd = torch.rsqrt(run_var + 1e-5)
        res2 = residual.contiguous().view(-1, Cc) if residual is not None else None
        C().batchnorm_fwd(x2, y, gamma, beta, mean, rstd, run_mean if training else None,
                          run_var if training else None, rows, Cc, 1e-5, 0.1, training, relu, res2)
        ctx.save_for_backward(x2, y, gamma, mean, rstd)
        ctx.gg, ctx.gb, ctx.relu, ctx.has_res, ctx.shape = ggamma, gbeta, relu, residual is not None, shape
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        x2, y, gamma, mean, rstd = ctx.saved_tensors
        Cc = x2.shape[1]
        dy2 = dy.contiguous().view(-1, Cc)
        dx = torch.empty_like(x2)
        dres = torch.empty_like(x2) if ctx.has_res else None
        gg = ctx.gg if ctx.gg is not None else torch.zeros(Cc, device=dy.device)
        gb = ctx.gb if ctx.gb is not None else torch.zeros(Cc, device=dy.device)
        C().batchnorm_bwd(dy2, x2, y, gamma, mean, rstd, dx, gg, gb, dres, x2.shape[0], Cc, ctx.relu)
        return (dx.view(ctx.shape), None, None, None, None, None, None, None, None,
                dres.view(ctx.shape) if dres is not None else None)


def batchnorm(x, gamma, beta, ggamma, gbeta, run_mean, run_var, training=True, relu=False,
              residual=None):
    return BatchNormFn.apply(x, gamma, beta, ggamma, gbeta, run_mean, run_var, training, relu, residual)


class MaxPoolFn(Function):
    @staticmethod
    def forward(ctx, x, k, stride, pad):
        x = x.contiguous()
        N, H, W, Cc = x.shape
        OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        y = torch.empty(N, OH, OW, Cc, device=x.device, dtype=BF)
        idx = torch.empty(N, OH, OW, Cc, device=x.device, dtype=torch.int32)
        C().maxpool_fwd(x, y, idx, N, Cc, H, W, k, stride, pad, OH, OW)
        ctx.save_for_backward(idx)
        ctx.in_shape = (N, H, W, Cc)
        return y

    @staticmethod
    def backward(ctx, dy):
        idx, = ctx.saved_tensors
        N, H, W, Cc = ctx.in_shape
        dxf = torch.zeros(N, H, W, Cc, device=dy.device, dtype=torch.float32)
        C().maxpool_bwd(dy.contiguous(), idx, dxf, idx.numel() // N, H * W * Cc)
        dx = torch.empty(N, H, W, Cc, device=dy.device, dtype=BF)
        C().cast_f32_to_bf16(dxf.view(-1), dx.view(-1))
        return dx, None, None, None


def maxpool2d(x, k=2, stride=2, pad=0):
    return MaxPoolFn.apply(x, k, stride, pad)


class GlobalAvgPoolFn(Function):
    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        N, H, W, Cc = x.shape
        y = torch.empty(N, Cc, device=x.device, dtype=BF)
        C().avgpool_fwd(x, y, N, H * W, Cc)
        ctx.shape = (N, H, W, Cc)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, H, W, Cc = ctx.shape
        dx = torch.empty(N, H, W, Cc, device=dy.device, dtype=BF)
        C().avgpool_bwd(dy.contiguous(), dx, N, H * W, Cc)
        return dx


def global_avgpool(x):
    return GlobalAvgPoolFn.apply(x)


class AddFn(Function):
    @staticmethod
    def forward(ctx, a, b):
        out = torch.empty_like(a)
        C().add_bf16(a.contiguous(), b.contiguous(), out)
        return out

    @staticmethod
    def backward(ctx, g):
        return g, g


def add(a, b):
    return AddFn.apply(a, b)


class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, ggamma, gbeta):
        x = x.contiguous()
        rows, Cc = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        C().layernorm_fwd(x, y, gamma, beta, mean, rstd, rows, Cc, 1e-12)
        ctx.save_for_backward(x, gamma, mean, rstd)
        ctx.gg, ctx.gb = ggamma, gbeta
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        rows, Cc = x.shape
        dx = torch.empty_like(x)
        gg = ctx.gg if ctx.gg is not None else torch.zeros(Cc, device=dy.device)
        gb = ctx.gb if ctx.gb is not None else torch.zeros(Cc, device=dy.device)
        C().layernorm_bwd(dy.contiguous(), x, gamma, mean, rstd, dx, gg, gb, rows, Cc)
        return dx, None, None, None, None


def layernorm(x, gamma, beta, ggamma=None, gbeta=None):
    return LayerNormFn.apply(x, gamma, beta, ggamma, gbeta)


class EmbeddingFn(Function):
    @staticmethod
    def forward(ctx, anchor, ids, table, pos, gtable, gpos, seq):
        rows, Cc = ids.numel(), table.shape[1]
        out = torch.empty(rows, Cc, device=table.device, dtype=BF)
        C().embedding_fwd(ids, table, pos, out, rows, seq, Cc)
        ctx.save_for_backward(ids)
        ctx.gt, ctx.gp, ctx.seq, ctx.C = gtable, gpos, seq, Cc
        return out

    @staticmethod
    def backward(ctx, dy):
        ids, = ctx.saved_tensors
        if ctx.gt is not None:
            C().embedding_bwd(ids, dy.contiguous(), ctx.gt, ctx.gp, ids.numel(), ctx.seq, ctx.C)
        return None, None, None, None, None, None, None


def embedding(ids, table, pos, gtable, gpos, seq):
    # `anchor` makes the output join the autograd graph even though no input is a leaf
    anchor = torch.zeros(1, device=table.device, requires_grad=gtable is not None)
    return EmbeddingFn.apply(anchor, ids, table, pos, gtable, gpos, seq)


class FusedAttentionFn(Function):
    """Multi-head self-attention core on q, k, v of shape [B*S, H*64], seq_len 128: ONE kernel per
    direction, one CTA per (batch, head) -- QK^T, softmax, PV (and in backward the five GEMMs of
    dQ / dK / dV) on tcgen05 with the S x S matrix held in TMEM / smem only, heads addressed as TMA
    boxes of the projection outputs so no transpose exists (csrc/kernels/attn_sm100.cu)."""

    @staticmethod
    def forward(ctx, q, k, v, B, S, H):
        D = q.shape[1] // H
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        out = torch.empty_like(q)
        lse = torch.empty(B * H * S, device=q.device, dtype=torch.float32)
        C().attention_fwd(q, k, v, out, lse, B, S, H, 1.0 / (D ** 0.5))
        ctx.save_for_backward(q, k, v, out, lse)
        ctx.dims = (B, S, H, D)
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, out, lse = ctx.saved_tensors
        B, S, H, D = ctx.dims
        dout = dout.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
        C().attention_bwd(q, k, v, out, dout, lse, dq, dk, dv, B, S, H, 1.0 / (D ** 0.5))
        return dq, dk, dv, None, None, None


class AttentionFn(Function):
    """Unfused fallback for shapes the fused kernel does not cover (seq != 128 or head dim != 64):
    batched tcgen05 GEMMs + row-softmax kernel + head transposes."""

    @staticmethod
    def forward(ctx, q, k, v, B, S, H):
        D = q.shape[1] // H
        m = C()

        def heads(t):
            o = torch.empty(B * H, S, D, device=t.device, dtype=BF)
            m.transpose_0213(t.contiguous(), o, B, S, H, D)
            return o

        qh, kh, vh = heads(q), heads(k), heads(v)
        scores = G.gemm(qh, kh)                                   # [BH, S, S]
        probs = torch.empty_like(scores)
        m.softmax_fwd(scores, probs, B * H * S, S, 1.0 / (D ** 0.5))
        ctxh = G.gemm(probs, vh, b_mn=True)                       # [BH, S, D]
        out = torch.empty(B * S, H * D, device=q.device, dtype=BF)
        m.transpose_0213(ctxh, out, B, H, S, D)
        ctx.save_for_backward(qh, kh, vh, probs)
        ctx.dims = (B, S, H, D)
        return out

    @staticmethod
    def backward(ctx, dout):
        qh, kh, vh, probs = ctx.saved_tensors
        B, S, H, D = ctx.dims
        m = C()
        doh = torch.empty(B * H, S, D, device=dout.device, dtype=BF)
        m.transpose_0213(dout.contiguous(), doh, B, S, H, D)
        dv = G.gemm(probs, doh, a_mn=True, b_mn=True)             # P^T dO   [BH, S, D]
        dp = G.gemm(doh, vh)                                      # dO V^T   [BH, S, S]
        ds = torch.empty_like(dp)
        m.softmax_bwd(dp, probs, ds, B * H * S, S, 1.0 / (D ** 0.5))
        dq = G.gemm(ds, kh, b_mn=True)                            # dS K     [BH, S, D]
        dk = G.gemm(ds, qh, a_mn=True, b_mn=True)                 # dS^T Q   [BH, S, D]

        def unheads(t):
            o = torch.empty(B * S, H * D, device=t.device, dtype=BF)
            m.transpose_0213(t, o, B, H, S, D)
            return o

        return unheads(dq), unheads(dk), unheads(dv), None, None, None


def attention(q, k, v, B, S, H, fused: bool = True):
    if fused and S == 128 and q.shape[1] // H == 64 and q.shape[1] % 8 == 0:
        return FusedAttentionFn.apply(q, k, v, B, S, H)
    return AttentionFn.apply(q, k, v, B, S, H)
