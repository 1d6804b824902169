"""Checkpoint / resume.  In the reference "the ledger is the checkpoint": global model + epoch +
roles persist with the chain, clients are stateless and re-register.  Here a
checkpoint is the host ledger snapshot (blocks + live state, hash-verified on load), the global
model, the device ledger page, and the per-client optimizer state.

Works for ``FusedEngine`` and ``GenericFedEngine`` (same buffer names)."""
from __future__ import annotations

import io
import struct
from typing import Optional

import torch
import torch.distributed as dist

from .._native import ledger as _ledger


def save_checkpoint(path: str, eng) -> dict:
    """Collective (every rank calls it; each rank writes ``path`` with its rank suffix when
    world > 1).  Drains the device block ring first so the host chain is complete."""
    errs = eng.drain_blocks()
    if errs:
        raise RuntimeError(f"cannot checkpoint: host/device ledgers disagree: {errs[:2]}")
    torch.cuda.synchronize()
    st = eng.read_state()
    blob = dict(
        version=1, world=eng.world, rank=eng.rank, config=eng.cfg.to_json(), n_params=eng.n_params,
        epoch=st["epoch"], state_bytes=bytes(eng.state_bytes.cpu().numpy()),
        global_master=eng.global_master.detach().cpu().clone(),
        ledger=bytes(eng.host_ledger.snapshot()),
        opt_m=getattr(getattr(eng, "trainer", eng), "m", None),
        opt_v=getattr(getattr(eng, "trainer", eng), "v", None),
    )
    for k in ("opt_m", "opt_v"):
        if blob[k] is not None:
            blob[k] = blob[k].detach().cpu().clone()
    out = path if eng.world == 1 else f"{path}.rank{eng.rank}"
    torch.save(blob, out)
    return dict(path=out, epoch=st["epoch"], blocks=eng.host_ledger.n_blocks())


def load_checkpoint(path: str, eng) -> dict:
    """Restore into a freshly constructed engine of the same config/world.  Collective."""
    src = path if eng.world == 1 else f"{path}.rank{eng.rank}"
    blob = torch.load(src, map_location="cpu", weights_only=False)
    if blob["n_params"] != eng.n_params or blob["world"] != eng.world:
        raise ValueError("checkpoint does not match this engine (n_params / world)")
    L = _ledger()
    eng.host_ledger = L.Ledger.restore(blob["ledger"])  # verifies the hash chain
    epoch = blob["epoch"]
    g = blob["global_master"].to(eng.dev)
    for t in (eng.global_master, eng.work_master):
        t.copy_(g)
    for t in (eng.global_shadow, eng.work_shadow):
        t.copy_(g.to(torch.bfloat16))
    eng.state_bytes.copy_(torch.frombuffer(bytearray(blob["state_bytes"]), dtype=torch.uint8))
    # epoch-tagged flags: everything up to `epoch` has happened on every rank
    n_flags = eng.sz["FLAG_COUNT"]
    flags = eng.heap.view(eng.layout.offsets["flags"], [n_flags], torch.int32)
    flags.fill_(epoch)
    holder = getattr(eng, "trainer", eng)
    for k, name in (("opt_m", "m"), ("opt_v", "v")):
        if blob[k] is not None and getattr(holder, name, None) is not None:
            getattr(holder, name).copy_(blob[k].to(eng.dev))
    eng.drained = epoch
    torch.cuda.synchronize()
    if eng.world > 1:
        dist.barrier(group=eng.group)
    return dict(epoch=epoch, blocks=eng.host_ledger.n_blocks())
