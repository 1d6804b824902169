"""Symmetric HBM heap across the ranks of one NVSwitch box (Python face of
``csrc/runtime/symm_heap.cu``).

Every rank allocates the same number of bytes; after ``rendezvous`` each rank holds a mapped
pointer to every peer's allocation (kernels ld/st peer HBM directly over NVLink) and, when
the fabric supports NVLS, one multicast pointer whose stores land in all replicas.
``torch.distributed`` is only the bootstrap channel for the opaque handles -- it replaces
the reference's TLS "Channel"/p2p transport configuration (README.md:238-260), never the
data path.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from .._native import C


class SymmetricHeap:
    def __init__(self, nbytes: int, *, rank: int = 0, world: int = 1, device: int = 0,
                 group: Optional[dist.ProcessGroup] = None, mode: str = "auto",
                 want_multicast: bool = True):
        self.rank, self.world, self.device = rank, world, device
        self.group = group
        self.multicast_error = ""
        mod = C()
        if world == 1:
            mode = "local"
        elif mode == "auto":
            mode = os.environ.get("BFLC_SYMM_MODE", "vmm")
        self.mode = mode
        self._heap = None
        if mode == "vmm":
            try:
                self._heap = mod.SymmHeap(nbytes, rank, world, device, "vmm")
                ok = 1
            except RuntimeError as e:  # VMM export refused in this container -> IPC
                self.multicast_error = f"vmm alloc failed: {e}"
                ok = 0
            if world > 1:
                flags = [None] * world
                dist.all_gather_object(flags, ok, group=group)
                if not all(flags):
                    self._heap = None
                    self.mode = mode = "ipc"
        if self._heap is None:
            self._heap = mod.SymmHeap(nbytes, rank, world, device, mode)
        self.nbytes = self._heap.bytes()
        self.mc_ptr = 0
        if world > 1:
            self._rendezvous(want_multicast)
        self.peer_ptrs: List[int] = [self._heap.peer_ptr(r) for r in range(world)]
        self.local_ptr = self._heap.local_ptr()

    # ------------------------------------------------------------------ bootstrap
    def _rendezvous(self, want_multicast: bool):
        blobs: List[Optional[bytes]] = [None] * self.world
        dist.all_gather_object(blobs, self._heap.export_handle(), group=self.group)
        try:
            self._heap.import_handles(blobs)
            ok = 1
        except RuntimeError as e:
            self.multicast_error = f"import failed: {e}"
            ok = 0
        oks = [None] * self.world
        dist.all_gather_object(oks, ok, group=self.group)
        if not all(oks):
            if self.mode == "vmm":
                # fall back collectively to CUDA IPC (e.g. pidfd_getfd blocked by seccomp)
                self.mode = "ipc"
                self._heap = C().SymmHeap(self.nbytes, self.rank, self.world, self.device, "ipc")
                self.nbytes = self._heap.bytes()
                blobs = [None] * self.world
                dist.all_gather_object(blobs, self._heap.export_handle(), group=self.group)
                self._heap.import_handles(blobs)
            else:
                raise RuntimeError("symmetric heap rendezvous failed: " + self.multicast_error)
        if self.mode == "vmm" and want_multicast:
            self._setup_multicast()
        dist.barrier(group=self.group)

    def _setup_multicast(self):
        blob = self._heap.mc_create_and_export() if self.rank == 0 else b""
        box = [blob]
        dist.broadcast_object_list(box, src=0, group=self.group)
        blob = box[0]
        ok = 1 if (blob and self._heap.mc_import_and_add(blob)) else 0
        oks = [None] * self.world
        dist.all_gather_object(oks, ok, group=self.group)  # barrier: all devices added
        if not all(oks):
            self.multicast_error = self._heap.last_error() or "multicast unsupported"
            return
        ok = 1 if self._heap.mc_bind_and_map() else 0
        dist.all_gather_object(oks, ok, group=self.group)
        if all(oks):
            self.mc_ptr = self._heap.mc_ptr()
        else:
            self.multicast_error = self._heap.last_error() or "multicast bind failed"

    # ------------------------------------------------------------------ views
    def view(self, offset: int, shape: Sequence[int], dtype: torch.dtype,
             rank: Optional[int] = None) -> torch.Tensor:
        """Tensor aliasing ``shape`` elements at byte ``offset`` of rank ``rank``'s heap (own
        heap by default).  A peer view is ordinary device memory to every kernel."""
        base = self.local_ptr if rank is None else self.peer_ptrs[rank]
        return C().tensor_from_ptr(base + offset, list(shape), dtype, self.device)

    def mc_view(self, offset: int, shape: Sequence[int], dtype: torch.dtype) -> torch.Tensor:
        assert self.mc_ptr, "multicast mapping unavailable: " + self.multicast_error
        return C().tensor_from_ptr(self.mc_ptr + offset, list(shape), dtype, self.device)

    @property
    def has_multicast(self) -> bool:
        return self.mc_ptr != 0

    def describe(self) -> dict:
        return dict(mode=self.mode, bytes=self.nbytes, world=self.world,
                    multicast=self.has_multicast, multicast_error=self.multicast_error)
