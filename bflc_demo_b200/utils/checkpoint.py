"""Checkpoint / resume.  In the reference "the ledger is the checkpoint": global model + epoch +
roles persist with the chain, clients are stateless and re-register.  Here a
checkpoint is the host ledger snapshot (blocks + live state, hash-verified on load), the global
model, the device ledger page, and the per-client optimizer state.

Works for ``FusedEngine`` and ``GenericFedEngine`` (same buffer names)."""
from __future__ import annotations

import struct

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
    # Only tensors / str / int: the file loads with torch.load(weights_only=True) (no pickled
    # objects).  Byte strings are stored as uint8 tensors.
    u8 = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).clone()   # noqa: E731
    opt_total, round_seq = _plan_counters(eng)
    blob = dict(
        version=2, world=eng.world, rank=eng.rank, config=eng.cfg.to_json(), n_params=eng.n_params,
        epoch=st["epoch"], state_bytes=eng.state_bytes.cpu().clone(),
        global_master=eng.global_master.detach().cpu().clone(),
        ledger=u8(eng.host_ledger.snapshot()),
        # Adam: moments AND the step count t the bias correction is computed from (the plan
        # page's running total of optimizer steps on this rank)
        opt_total=opt_total, round_seq=round_seq,
    )
    holder = getattr(eng, "trainer", eng)
    for k, name in (("opt_m", "m"), ("opt_v", "v")):
        t = getattr(holder, name, None)
        if t is not None:
            blob[k] = t.detach().cpu().clone()
    out = path if eng.world == 1 else f"{path}.rank{eng.rank}"
    torch.save(blob, out)
    return dict(path=out, epoch=st["epoch"], blocks=eng.host_ledger.n_blocks())


def _plan_view(eng):
    sz = eng.sz
    return eng.heap.view(eng.layout.offsets["plan"], [sz["RoundPlan"]], torch.uint8)


def _plan_counters(eng):
    raw = bytes(_plan_view(eng).cpu().numpy())
    opt_total, = struct.unpack_from("<i", raw, eng.sz["plan_opt_total_off"])
    round_seq, = struct.unpack_from("<I", raw, eng.sz["plan_round_seq_off"])
    return int(opt_total), int(round_seq)


def _set_plan_counters(eng, opt_total: int, round_seq: int):
    view = _plan_view(eng)
    raw = bytearray(bytes(view.cpu().numpy()))
    struct.pack_into("<i", raw, eng.sz["plan_opt_total_off"], int(opt_total))
    struct.pack_into("<i", raw, eng.sz["plan_opt_step_off"], int(opt_total))
    struct.pack_into("<I", raw, eng.sz["plan_round_seq_off"], int(round_seq))
    view.copy_(torch.frombuffer(raw, dtype=torch.uint8))


def load_checkpoint(path: str, eng) -> dict:
    """Restore into a freshly constructed engine of the same config/world.  Collective."""
    src = path if eng.world == 1 else f"{path}.rank{eng.rank}"
    blob = torch.load(src, map_location="cpu", weights_only=True)
    if blob["n_params"] != eng.n_params or blob["world"] != eng.world:
        raise ValueError("checkpoint does not match this engine (n_params / world)")
    L = _ledger()
    eng.host_ledger = L.Ledger.restore(bytes(blob["ledger"].numpy()))  # verifies the hash chain
    epoch = blob["epoch"]
    g = blob["global_master"].to(eng.dev)
    for t in (eng.global_master, eng.work_master):
        t.copy_(g)
    for t in (eng.global_shadow, eng.work_shadow):
        t.copy_(g.to(torch.bfloat16))
    eng.state_bytes.copy_(blob["state_bytes"])
    # epoch-tagged flags: everything up to `epoch` has happened on every rank
    n_flags = eng.sz["FLAG_COUNT"]
    flags = eng.heap.view(eng.layout.offsets["flags"], [n_flags], torch.int32)
    flags.fill_(epoch)
    holder = getattr(eng, "trainer", eng)
    for k, name in (("opt_m", "m"), ("opt_v", "v")):
        if blob.get(k) is not None and getattr(holder, name, None) is not None:
            getattr(holder, name).copy_(blob[k].to(eng.dev))
    # Adam's t continues where the saved run stopped, whatever warm-up rounds this engine ran
    # before the restore (FusedEngine.capture() runs one); the input-pipeline generation keeps
    # growing (tags only ever increase), so round_seq is restored only if it moves forward
    _, cur_seq = _plan_counters(eng)
    _set_plan_counters(eng, blob.get("opt_total", 0), max(cur_seq, int(blob.get("round_seq", 0))))
    eng.drained = epoch
    eng._rounds = epoch
    # host-side caches of the ledger page are stale now: the e2e path re-learns the epoch with one
    # synchronous read, the generic engine re-reads its role table
    if hasattr(eng, "_epoch_known"):
        eng._epoch_known = None
    if hasattr(eng, "_st"):
        eng._st = None
    torch.cuda.synchronize()
    if eng.world > 1:
        dist.barrier(group=eng.group)
    return dict(epoch=epoch, blocks=eng.host_ledger.n_blocks())
