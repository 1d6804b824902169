"""Python face of the tcgen05 GEMM (csrc/kernels/gemm_sm100.cu).

``D[b] = alpha * A[b] @ B[b]^T`` with A given as ``[M, K]`` (K-major) or ``[K, M]``
(``a_mn=True``) and B as ``[N, K]`` or ``[K, N]`` (``b_mn=True``); the transposed forms are
consumed directly through MN-major UMMA descriptors, never materialised.

Reference parity: K1 ``tf.matmul(x, W) + b`` (python-sdk/main.py:120,180,293).
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from .._native import C

EPI_GENERIC, EPI_XENT, EPI_ARGMAX = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
_DT = {torch.float32: 0, torch.bfloat16: 1, torch.float8_e4m3fn: 2}
_GEMM2 = os.environ.get("BFLC_GEMM2", "1") != "0"   # route big K-major GEMMs to the CTA-pair kernel


def _mat_dims(t: torch.Tensor, mn: bool):
    """rows(M|N), K, ld, batch, batch_stride of an operand (2-D or 3-D, last dim contiguous)."""
    assert t.stride(-1) == 1, "operand rows must be contiguous"
    if t.dim() == 2:
        r, c = t.shape
        batch, bs = 1, 0
    else:
        batch, r, c = t.shape
        bs = t.stride(0) if batch > 1 else 0
    ld = t.stride(-2)
    if mn:
        return c, r, ld, batch, bs  # memory is [K][MN]
    return r, c, ld, batch, bs


def gemm(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, *,
         a_mn: bool = False, b_mn: bool = False, out_dtype: torch.dtype = torch.bfloat16,
         alpha: float = 1.0, bias: Optional[torch.Tensor] = None, act: int = ACT_NONE,
         aux_out: Optional[torch.Tensor] = None, aux_in: Optional[torch.Tensor] = None,
         act_bwd: int = 0, colsum: Optional[torch.Tensor] = None, split_k: int = 1,
         accumulate: bool = False, n_valid: Optional[int] = None,
         b_maps: Optional[torch.Tensor] = None, dyn_ptr: int = 0, batch: Optional[int] = None,
         dbg=(0, 0, 0, 0)) -> torch.Tensor:
    is_fp8 = a.dtype == torch.float8_e4m3fn
    M, K, lda, ba, a_bs = _mat_dims(a, a_mn)
    N, Kb, ldb, bb, b_bs = _mat_dims(b, b_mn)
    assert K == Kb, f"K mismatch {K} vs {Kb}"
    if n_valid is not None:
        N = n_valid
    nb = batch if batch is not None else max(ba, bb)
    if out is None:
        shape = (M, N) if nb == 1 and a.dim() == 2 else (nb, M, N)
        if split_k > 1:
            out = torch.zeros(shape, device=a.device, dtype=torch.float32)
        else:
            out = torch.empty(shape, device=a.device, dtype=out_dtype)
    ldd = out.stride(-2)
    d_bs = out.stride(0) if out.dim() == 3 else 0
    # Large plain K-major problems go to the CTA-pair kernel (cta_group::2, 256x256 tiles):
    # the 1-CTA 128x256 tile is bound by the L2->SM feed rate, the pair moves 2/3 of the bytes.
    if (_GEMM2 and not is_fp8 and not a_mn and not b_mn and nb == 1 and a.dim() == 2 and split_k == 1
            and not accumulate and aux_out is None and aux_in is None and colsum is None
            and b_maps is None and dyn_ptr == 0 and out.dtype in (torch.float32, torch.bfloat16)
            and ldd % 4 == 0 and M >= 512 and N >= 512 and M * N >= (1 << 22)
            and C().current_predicate_is_null()):
        C().gemm2(a, b, out, M, N, K, lda, ldb, alpha, bias, act)
        return out
    C().gemm(a, b, out, M, N, K, nb, lda, ldb, a_bs, b_bs, a_mn, b_mn, is_fp8, EPI_GENERIC,
             _DT[out.dtype], ldd, d_bs, alpha, bias, act, aux_out, aux_in, act_bwd, colsum,
             split_k, accumulate, None, 0, 1.0, None, None, b_maps, None, *dbg, dyn_ptr)
    return out


def gemm_xent(a: torch.Tensor, b: torch.Tensor, labels: torch.Tensor, *, n_classes: int,
              bias: Optional[torch.Tensor], dlogits: Optional[torch.Tensor], grad_scale: float,
              loss_sum: torch.Tensor, correct: Optional[torch.Tensor] = None,
              colsum: Optional[torch.Tensor] = None, b_mn: bool = False, alpha: float = 1.0):
    """logits = A @ B^T + bias fused with softmax-cross-entropy: writes dlogits (bf16, padded
    to its row stride), accumulates the loss sum and #correct.  K2 + K6 of the reference
    (python-sdk/main.py:123, 182-183)."""
    is_fp8 = a.dtype == torch.float8_e4m3fn
    M, K, lda, ba, a_bs = _mat_dims(a, False)
    N, Kb, ldb, bb, b_bs = _mat_dims(b, b_mn)
    assert K == Kb
    ldd = dlogits.stride(-2) if dlogits is not None else 0
    C().gemm(a, b, dlogits, M, n_classes, K, 1, lda, ldb, 0, 0, False, b_mn, is_fp8, EPI_XENT, 1,
             ldd, 0, alpha, bias, 0, None, None, 0, colsum, 1, False, labels, 0, grad_scale,
             loss_sum, correct, None, None, 0, 0, 0, 0)


def gemm_argmax_acc(a: torch.Tensor, b: torch.Tensor, labels: torch.Tensor, correct: torch.Tensor,
                    *, n_classes: int, bias: Optional[torch.Tensor] = None, batch: int = 1,
                    a_bs: int = 0, b_bs: int = 0, labels_bs: int = 0,
                    b_maps: Optional[torch.Tensor] = None, dyn_ptr: int = 0, alpha: float = 1.0):
    """#(argmax(A @ B^T + bias) == label) per batch entry -> correct[b]  (committee score)."""
    is_fp8 = a.dtype == torch.float8_e4m3fn
    M, K = a.shape[-2], a.shape[-1]
    C().gemm(a, b, None, M, n_classes, K, batch, a.stride(-2), b.stride(-2), a_bs, b_bs, False,
             False, is_fp8, EPI_ARGMAX, 1, 0, 0, alpha, bias, 0, None, None, 0, None, 1, False,
             labels, labels_bs, 1.0, None, correct, b_maps, None, 0, 0, 0, 0, dyn_ptr)


def gemm_2cta(a: torch.Tensor, b: torch.Tensor, out: Optional[torch.Tensor] = None, *,
              out_dtype: torch.dtype = torch.bfloat16, alpha: float = 1.0,
              bias: Optional[torch.Tensor] = None, act: int = ACT_NONE) -> torch.Tensor:
    """Large-problem path: CTA pairs (``tcgen05.mma.cta_group::2``, UMMA M = 256) computing
    256 x 256 tiles; ``a`` [M, K] and ``b`` [N, K] bf16, K-major (csrc/kernels/gemm2_sm100.cu)."""
    M, K = a.shape
    N = b.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    C().gemm2(a, b, out, M, N, K, a.stride(0), b.stride(0), alpha, bias, act)
    return out
