"""Flat parameter buffers: every model's tensors are views into ONE contiguous fp32 master
buffer (plus a bf16 shadow the tensor cores read and a fp32 gradient buffer), so that

* the optimizer is one fused kernel over the whole model,
* an "update" in the protocol sense (reference ``LocalUpdate.delta_model``,
  CommitteePrecompiled.h:82-107) is a single address range in the symmetric heap that peers
  can read with P2P loads / TMA, and
* FedAvg is one pass over one range.

The reference's ``Model`` is two nested ``std::vector<float>`` (``ser_W[5][2]``, ``ser_b[2]``,
H:24-52) serialised to JSON on every access; this is its binary, zero-copy replacement.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch


def _up8(n: int) -> int:
    return (n + 7) // 8 * 8


@dataclass
class ParamEntry:
    name: str
    shape: Tuple[int, ...]
    offset: int  # elements, multiple of 8 (16-byte aligned in bf16 -> valid TMA base)
    numel: int


class ParamSpec:
    def __init__(self, entries: Sequence[Tuple[str, Sequence[int]]]):
        self.entries: List[ParamEntry] = []
        cur = 0
        for name, shape in entries:
            n = 1
            for s in shape:
                n *= int(s)
            self.entries.append(ParamEntry(name, tuple(int(s) for s in shape), cur, n))
            cur = _up8(cur + n)
        self.total = _up8(cur)
        self.by_name: Dict[str, ParamEntry] = {e.name: e for e in self.entries}

    def views(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        assert flat.numel() >= self.total
        return {e.name: flat[e.offset:e.offset + e.numel].view(e.shape) for e in self.entries}

    def offset(self, name: str) -> int:
        return self.by_name[name].offset

    def init_(self, flat: torch.Tensor, seed: int = 0) -> None:
        """Deterministic init (same on every rank -> identical genesis global model):
        matrices ~ U(-1/sqrt(fan_in), 1/sqrt(fan_in)); vectors named *gamma*/*weight_ln* = 1,
        other vectors = 0.  The reference's genesis model is all zeros (C:325-327), which is
        only workable for its single linear layer."""
        g = torch.Generator(device="cpu").manual_seed(seed)
        flat.zero_()
        v = self.views(flat)
        for e in self.entries:
            t = v[e.name]
            if len(e.shape) >= 2:
                fan_in = 1
                for s in e.shape[1:]:
                    fan_in *= s
                bound = 1.0 / (fan_in ** 0.5)
                t.copy_(((torch.rand(e.shape, generator=g) * 2 - 1) * bound).to(t.device))
            elif "gamma" in e.name:
                t.fill_(1.0)
