"""Block-scaled fp8 (OCP MXFP8) operands and GEMM (csrc/kernels/gemm_mx8_sm100.cu).

An :class:`MX8` holds e4m3 elements ``q [R, ld]`` plus one UE8M0 scale per 32 K-elements in
the chunk layout the tensor core consumes (``tcgen05.cp`` copies a chunk into TMEM as is).
``gemm_mx8(a, b)`` = ``(a.q * a.scale) @ (b.q * b.scale)^T`` on
``tcgen05.mma.kind::mxf8f6f4.block_scale``.

Reference parity: the reference's dense layers (python-sdk/main.py:120-123) in the precision
BASELINE.json names for the MLP / LeNet-5 configs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch

from .._native import C

SF_CHUNK = 512


@dataclass
class MX8:
    q: torch.Tensor          # float8_e4m3fn [R, ld]  (ld = K rounded up to 16)
    sf: torch.Tensor         # uint8 scale chunks [row_block][k_block][512]
    rows: int
    K: int

    def dequantize(self) -> torch.Tensor:
        """fp32 [rows, K] (test / debug path, plain PyTorch)."""
        kb = (self.K + 127) // 128
        rb = self.sf.numel() // (kb * SF_CHUNK)
        s = self.sf.view(rb, kb, 32, 4, 4).permute(0, 3, 2, 1, 4)      # [rb, r1, r0, kb, kk]
        s = s.reshape(rb * 128, kb * 4)[: self.rows]
        scale = torch.exp2(s.float() - 127.0).repeat_interleave(32, dim=1)[:, : self.K]
        return self.q[:, : self.K].float() * scale


def quantize_mx8_reference(x: torch.Tensor, *, in_scale: float = 1.0) -> MX8:
    """Plain-PyTorch (CPU or GPU) specification of ``k_quantize_mx8``: per (row, 32-element K
    group) the UE8M0 exponent ``e = ceil(log2(amax / 448))`` (clamped to [1, 254] biased), elements
    ``e4m3(x * 2^-e)`` saturating, scales stored in the tensor-core chunk layout -- per (128-row
    block, 128-K block) 512 bytes, byte ``[r % 32][r // 32][k // 32]``.  Rows are padded to 256,
    K groups to a multiple of 4; padding carries scale 1.0 (0x7F)."""
    assert x.dim() == 2
    R, K = x.shape
    xf = x.float() * in_scale
    kb = (K + 127) // 128
    G = kb * 4
    pad = torch.zeros(R, G * 32, dtype=torch.float32, device=x.device)
    pad[:, :K] = xf
    g = pad.view(R, G, 32)
    amax = g.abs().amax(-1)
    e = torch.full_like(amax, 127.0)
    nz = amax > 0
    e[nz] = torch.ceil(torch.log2(amax[nz] / 448.0)).clamp(-126, 127) + 127.0
    n_real = (K + 31) // 32
    e[:, n_real:] = 127.0                                  # groups entirely beyond K
    scale = torch.exp2(e - 127.0)
    q = (g / scale.unsqueeze(-1)).clamp(-448.0, 448.0).view(R, G * 32)
    ld = (K + 15) // 16 * 16
    qq = torch.zeros(R, ld, dtype=torch.float8_e4m3fn, device=x.device)
    qq[:, :K] = q[:, :K].to(torch.float8_e4m3fn)
    rb = (R + 255) // 256 * 2
    sf_rows = torch.full((rb * 128, G), 127, dtype=torch.uint8, device=x.device)
    sf_rows[:R] = e.to(torch.uint8)
    # [rb, r1, r0, kb, kk] -> chunk bytes [rb, kb, r0, r1, kk]
    sf = sf_rows.view(rb, 4, 32, kb, 4).permute(0, 3, 2, 1, 4).contiguous().view(-1)
    return MX8(qq, sf, R, K)


def quantize_mx8(x: torch.Tensor, *, in_scale: float = 1.0, out: Optional[MX8] = None) -> MX8:
    """x [R, K] (f32 / bf16 / u8) * in_scale -> MX8 (scales along K)."""
    assert x.dim() == 2 and x.stride(1) == 1
    R, K = x.shape
    if out is None:
        ld = (K + 15) // 16 * 16
        q = torch.zeros(R, ld, device=x.device, dtype=torch.float8_e4m3fn)
        sf = torch.empty(C().mx8_sf_bytes(R, K), device=x.device, dtype=torch.uint8)
        out = MX8(q, sf, R, K)
    C().quantize_mx8(x, out.q, out.sf, R, K, in_scale)
    return out


def gemm_mx8(a: MX8, b: MX8, out: Optional[torch.Tensor] = None, *,
             out_dtype: torch.dtype = torch.bfloat16, alpha: float = 1.0,
             bias: Optional[torch.Tensor] = None, act: int = 0) -> torch.Tensor:
    assert a.K == b.K, f"K mismatch {a.K} vs {b.K}"
    M, N = a.rows, b.rows
    if out is None:
        ldd = (N + 3) // 4 * 4
        buf = torch.empty(M, ldd, device=a.q.device, dtype=out_dtype)
        out = buf[:, :N] if ldd != N else buf
    C().gemm_mx8(a.q, a.sf, b.q, b.sf, out, M, N, a.K, alpha, bias, act)
    return out
