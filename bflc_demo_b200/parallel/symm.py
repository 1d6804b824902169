"""Symmetric HBM heap across the ranks of one NVSwitch box (Python face of
``csrc/runtime/symm_heap.cu``).

Every rank allocates the same number of bytes; after the rendezvous each rank holds a mapped
pointer to every peer's allocation (kernels ld/st peer HBM directly over NVLink) and, when
the fabric supports NVLS, one multicast pointer whose stores land in all replicas.
``torch.distributed`` is only the bootstrap channel for names / opaque handles -- it replaces
the reference's TLS "Channel"/p2p transport configuration (README.md:238-260), never the
data path.

Substrates, tried in this order (``mode="auto"``):
  vmm   cuMemCreate + POSIX fds passed over abstract unix sockets (SCM_RIGHTS); the only
        substrate that can carry an NVLS multicast mapping
  ipc   cudaMalloc + cudaIpc handles (no multicast)
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
        self.notes: List[str] = []
        self.mc_ptr = 0
        mod = C()
        if world == 1:
            self.mode = "local"
            self._heap = mod.SymmHeap(nbytes, rank, world, device, "local")
        else:
            if mode == "auto":
                mode = os.environ.get("BFLC_SYMM_MODE", "vmm")
            self._heap = None
            if mode == "vmm":
                self._try_vmm(nbytes, want_multicast)
            if self._heap is None:
                self._open_ipc(nbytes)
            dist.barrier(group=group)
        self.nbytes = self._heap.bytes()
        self.peer_ptrs: List[int] = [self._heap.peer_ptr(r) for r in range(world)]
        self.local_ptr = self._heap.local_ptr()

    # ------------------------------------------------------------------ helpers
    def _all_ok(self, ok: bool) -> bool:
        flags: List[Optional[int]] = [None] * self.world
        dist.all_gather_object(flags, 1 if ok else 0, group=self.group)
        return all(flags)

    def _gather(self, obj) -> list:
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    # ------------------------------------------------------------------ substrates
    def _try_vmm(self, nbytes: int, want_multicast: bool):
        mod = C()
        heap, err = None, ""
        try:
            heap = mod.SymmHeap(nbytes, self.rank, self.world, self.device, "vmm")
        except RuntimeError as e:
            err = f"vmm alloc: {e}"
        if not self._all_ok(heap is not None):
            self.notes.append(err or "vmm alloc failed on a peer")
            return
        # unique tag: rank 0's pid + a counter so several heaps can coexist
        tag_box = [f"{os.getpid()}_{id(self) & 0xffffff}"]
        dist.broadcast_object_list(tag_box, src=0, group=self.group)
        names, err = None, ""
        try:
            names = self._gather(heap.fd_listen(tag_box[0]))  # gather doubles as the barrier
            heap.import_via_sockets(names)
            ok = True
        except RuntimeError as e:
            ok, err = False, f"vmm socket import: {e}"
        if not self._all_ok(ok):
            self.notes.append(err or "vmm import failed on a peer")
            del heap
            return
        self._heap, self.mode = heap, "vmm"
        if want_multicast:
            self._setup_multicast(names)

    def _setup_multicast(self, names):
        if self.rank == 0:
            blob = self._heap.mc_create_and_export()
            if not blob:
                self.multicast_error = self._heap.last_error() or "multicast unsupported"
        # every rank joins the exchange even if rank 0 failed (it then sends "no fd")
        try:
            ok = bool(self._heap.mc_import_via_sockets(names))
        except RuntimeError as e:
            ok = False
            self.multicast_error = str(e)
        if not self._all_ok(ok):          # also the "all devices added" barrier
            self.multicast_error = self.multicast_error or self._heap.last_error() or \
                "multicast add-device failed on a peer"
            return
        ok = bool(self._heap.mc_bind_and_map())
        if self._all_ok(ok):
            self.mc_ptr = self._heap.mc_ptr()
        else:
            self.multicast_error = self._heap.last_error() or "multicast bind failed on a peer"

    def _open_ipc(self, nbytes: int):
        self.mode = "ipc"
        self._heap = C().SymmHeap(nbytes, self.rank, self.world, self.device, "ipc")
        blobs = self._gather(self._heap.export_handle())
        self._heap.import_handles(blobs)

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
                    multicast=self.has_multicast, multicast_error=self.multicast_error,
                    notes=self.notes)
