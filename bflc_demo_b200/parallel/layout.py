"""Byte layout of the symmetric heap (identical on every rank) -- Python mirror of
``bflc::HeapLayout`` (csrc/include/bflc_kernels.h).  The regions replace the seven JSON
strings the reference keeps in one KV table (CommitteePrecompiled.cpp:32-44): flags and
RoundState are the "epoch/roles/counters" keys, the score matrix and upload buffers are
``local_scores`` / ``local_updates``, ``global`` is ``global_model``."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List

from .._native import C


def _up(x: int, a: int) -> int:
    return (x + a - 1) // a * a


@dataclass
class HeapLayout:
    n_params: int                     # padded to a multiple of 8 elements
    ring_slots: int = 256
    extra_bytes: int = 0              # caller-owned scratch appended after the fixed regions
    offsets: Dict[str, int] = field(default_factory=dict)
    total_bytes: int = 0
    sizes: Dict[str, int] = field(default_factory=dict)

    def __post_init__(self):
        sz = C().struct_sizes()
        self.sizes = dict(sz)
        assert self.n_params % 8 == 0
        K = sz["kMaxRanks"]
        cur = 0

        def take(name: str, nbytes: int, align: int = 1024):
            nonlocal cur
            cur = _up(cur, align)
            self.offsets[name] = cur
            cur += nbytes

        take("flags", sz["FLAG_COUNT"] * 4)
        take("state", sz["RoundState"])
        take("plan", sz["RoundPlan"])
        take("scores", 2 * K * K * 4 + 2 * K * 8)   # score rows by parity + two-shot slice digests
        take("meta", 2 * K * sz["UploadMeta"])
        take("admit", 2 * sz["AdmitPage"])              # first-K-wins admission: ticket + slots, by parity
        take("ring", self.ring_slots * sz["BlockRecord"])
        f32, b16 = self.n_params * 4, self.n_params * 2
        take("work_master", f32, 4096)
        take("work_shadow", b16, 4096)
        take("upload_master0", f32, 4096)
        take("upload_master1", f32, 4096)
        take("upload_shadow0", b16, 4096)
        take("upload_shadow1", b16, 4096)
        take("global", f32, 4096)
        take("global_shadow", b16, 4096)
        take("extra", self.extra_bytes, 4096)
        self.total_bytes = _up(cur, 1 << 21)

    def fed_dict(self, rank: int, n_ranks: int, peer_bases: List[int], mc_base: int) -> dict:
        o = self.offsets
        return dict(rank=rank, n_ranks=n_ranks, peer_bases=list(peer_bases), mc_base=mc_base,
                    flags_off=o["flags"], state_off=o["state"], plan_off=o["plan"],
                    scores_off=o["scores"], meta_off=o["meta"],
                    work_master_off=o["work_master"], work_shadow_off=o["work_shadow"],
                    upload_master_off=[o["upload_master0"], o["upload_master1"]],
                    upload_shadow_off=[o["upload_shadow0"], o["upload_shadow1"]],
                    global_off=o["global"], global_shadow_off=o["global_shadow"],
                    ring_off=o["ring"], n_params=self.n_params, ring_slots=self.ring_slots,
                    admit_off=o["admit"])
