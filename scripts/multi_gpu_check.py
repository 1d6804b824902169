"""Multi-GPU protocol checks (run under torchrun, any world size 2..8). Rank 0 prints one
``RESULT {json}`` line.  Used by tests/test_gpu_multi.py and the gpurun scripts.

Checks:
  fused       MLP fused engine: rounds advance, host ledgers verify, replicas bit-identical,
              committee rotates, loss falls
  byzantine   one sign-flipping rank is never aggregated nor elected
  two_shot    two-shot (slice-reduce + publish) aggregation gives the same digest as one-shot
  multicast   the same through NVLS multimem stores when the heap has a multicast mapping
  generic     LeNet-5 through the model-agnostic engine (validation on peers' HBM)
  firstk      device-side first-K-wins admission (C:239-244): needed_updates = trainers - 1 and one
              artificially slow trainer -- every round completes with exactly K admitted, the
              straggler's update is dropped, the host ledger re-executes from the admitted mask
  fedavg      the aggregated global model is RIGHT, not just identical: after every round each rank
              recomputes sum_k w_k * upload_k (selected set + weights from the host ledger's block,
              uploads read out of the trainers' HBM, ascending rank order, fp32 fma) in PyTorch and
              compares it with the device result -- bf16 and fp8 engines
"""
import os as _os, sys as _sys
_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
import json
import os
import sys

import torch
import torch.distributed as dist

from bflc_demo_b200.config import FLConfig
from bflc_demo_b200.data.synthetic import cifar_like, femnist_like


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    which = sys.argv[1:] or ["fused", "byzantine", "two_shot", "generic"]
    out = {"world": world}

    def gather(x):
        box = [None] * world
        dist.all_gather_object(box, x)
        return box

    from bflc_demo_b200.engine.fused import FusedEngine

    def run_fused(rounds=6, **kw):
        kw.setdefault("dtype", os.environ.get("BFLC_CHECK_DTYPE", "bf16"))   # fp8: MXFP8 trainer + blobs
        cfg = FLConfig.for_world(world, hidden=256, batch_size=128, samples_per_client=512,
                                 learning_rate=0.05, **kw)
        shard = femnist_like(world, 512, seed=3, only=rank)[0]
        eng = FusedEngine(cfg, shard, rank=rank, world=world, device=lr)
        eng.capture()
        hist = [eng.run_round_e2e() for _ in range(rounds)]
        errs = eng.drain_blocks()
        st = eng.read_state()
        info = dict(epoch=st["epoch"], digest=st["model_digest"], errs=errs,
                    chain=eng.host_ledger.verify_chain(), blocks=eng.host_ledger.n_blocks(),
                    last_hash=eng.host_ledger.blocks()[-1]["hash"], loss=[h["global_loss"] for h in hist],
                    roles=[h["roles"] for h in hist], symm=eng.heap.describe())
        blocks = eng.host_ledger.blocks()
        torch.cuda.synchronize(); dist.barrier()   # nobody may still be reading my heap
        del eng
        torch.cuda.synchronize(); dist.barrier()
        return info, blocks

    if "fused" in which:
        info, blocks = run_fused()
        allinfo = gather(info)
        out["fused"] = dict(
            epochs=[i["epoch"] for i in allinfo], errs=sum((i["errs"] for i in allinfo), []),
            identical_digest=len({i["digest"] for i in allinfo}) == 1,
            identical_chain=len({i["last_hash"] for i in allinfo}) == 1,
            chain_ok=all(i["chain"] for i in allinfo), loss=info["loss"],
            committee_rotates=len({tuple(r) for r in info["roles"]}) > 1 if world > 2 else True,
            symm=info["symm"])
    if "byzantine" in which and world >= 4:
        byz = world - 1
        info, blocks = run_fused(rounds=6, byzantine_ranks=[byz], byzantine_scale=5.0)
        sel = [b["selected"] for b in blocks]
        elected = [b["role_after"][byz] for b in blocks]
        admitted = [byz in b["admitted"] for b in blocks]
        out["byzantine"] = dict(rank=byz, ever_selected=any(byz in s for s in sel),
                                ever_elected=any(e == 2 for e in elected),
                                times_admitted=sum(admitted),
                                median_of_byz=[b["median"][b["admitted"].index(byz)] for b in blocks if byz in b["admitted"]][:3],
                                median_best=[max(b["median"]) for b in blocks][:3])
    if "two_shot" in which:
        # (local training uses split-K atomics, so two separate runs are not bit-comparable;
        #  what must hold in every mode is that all replicas of one run are bit-identical)
        res = {}
        for name, kw in (("two_shot_p2p", dict(two_shot=True, use_multicast=False)),
                         ("two_shot_multicast", dict(two_shot=True, use_multicast=True))):
            r, _ = run_fused(rounds=4, **kw)
            g = gather(r)
            res[name] = dict(identical=len({i["digest"] for i in g}) == 1,
                             errs=sum((i["errs"] for i in g), []), chain_ok=all(i["chain"] for i in g),
                             loss=r["loss"], multicast=r["symm"]["multicast"],
                             multicast_error=r["symm"]["multicast_error"], notes=r["symm"]["notes"])
        out["two_shot"] = res
    if "firstk" in which and world >= 4:
        res = {}
        slow = world - 1
        for dt in ("bf16", "fp8"):
            base = FLConfig.for_world(world)
            k = base.n_trainers - 1
            cfg = FLConfig.for_world(world, needed_updates=k, hidden=256, batch_size=128,
                                     samples_per_client=512, learning_rate=0.05, dtype=dt,
                                     straggler_ranks=[slow], straggler_delay_us=400)
            shard = femnist_like(world, 512, seed=3, only=rank)[0]
            eng = FusedEngine(cfg, shard, rank=rank, world=world, device=lr)
            eng.capture()
            for _ in range(6):
                eng.run_round()
            errs = eng.drain_blocks()
            st = eng.read_state()
            blocks = eng.host_ledger.blocks()
            g = gather(dict(digest=st["model_digest"], errs=errs, epoch=st["epoch"]))
            res[dt] = dict(k=k, trainers=cfg.n_trainers, epoch=st["epoch"],
                           admitted_per_round=[len(b["admitted"]) for b in blocks],
                           slow_rank=slow, slow_was_trainer=sum(b["role_before"][slow] == 1 for b in blocks),
                           slow_admitted=sum(slow in b["admitted"] for b in blocks),
                           identical=len({i["digest"] for i in g}) == 1,
                           errs=sum((i["errs"] for i in g), []), chain_ok=eng.host_ledger.verify_chain())
            torch.cuda.synchronize(); dist.barrier()
            del eng
            torch.cuda.synchronize(); dist.barrier()
        out["firstk"] = res
    if "fedavg" in which:
        res = {}
        for dt in ("bf16", "fp8"):
            cfg = FLConfig.for_world(world, hidden=256, batch_size=128, samples_per_client=512,
                                     learning_rate=0.05, dtype=dt)
            shard = femnist_like(world, 512, seed=3, only=rank)[0]
            eng = FusedEngine(cfg, shard, rank=rank, world=world, device=lr)
            eng.capture()
            o, P = eng.layout.offsets, eng.n_params
            worst, exact, errs = 0.0, True, []
            for _ in range(4):
                eng.run_round()
                torch.cuda.synchronize(); dist.barrier()
                errs += eng.drain_blocks()
                blk = eng.host_ledger.blocks()[-1]
                par = blk["epoch"] & 1
                ref = torch.zeros(P, device="cuda", dtype=torch.float64)
                for t, w in zip(blk["selected"], blk["weight"]):
                    up = eng.heap.view(o[f"upload_master{par}"], [P], torch.float32, rank=t)
                    # fp32 fma(w, v, acc): the product is exact in fp64, one rounding back to fp32
                    ref = (ref + up.double() * float(w)).float().double()
                d = (eng.global_master.double() - ref).abs().max().item()
                worst = max(worst, d / max(ref.abs().max().item(), 1e-30))
                exact = exact and bool((eng.global_master.double() == ref).all())
                torch.cuda.synchronize(); dist.barrier()
            g = gather(dict(worst=worst, exact=exact, errs=errs, n_sel=len(blk["selected"])))
            res[dt] = dict(worst_rel=max(i["worst"] for i in g), bit_exact=all(i["exact"] for i in g),
                           errs=sum((i["errs"] for i in g), []), n_selected=g[0]["n_sel"])
            torch.cuda.synchronize(); dist.barrier()
            del eng
            torch.cuda.synchronize(); dist.barrier()
        out["fedavg"] = res
    if "generic" in which:
        from bflc_demo_b200.engine.generic import GenericFedEngine
        from bflc_demo_b200.models.nets import LeNet5
        cfg = FLConfig.for_world(world, batch_size=64, samples_per_client=256, learning_rate=0.05,
                                 model="lenet5", dataset="cifar10")
        shard = cifar_like(world, 256, seed=2)[rank]
        eng = GenericFedEngine(cfg, LeNet5(10), shard, rank=rank, world=world, device=lr)
        acc0 = eng.evaluate(shard)
        for _ in range(5):
            eng.run_round()
        errs = eng.drain_blocks()
        st = eng.read_state()
        g = gather(dict(digest=st["model_digest"], errs=errs, epoch=st["epoch"]))
        out["generic_lenet5"] = dict(epoch=st["epoch"], acc_before=acc0, acc_after=eng.evaluate(shard),
                                     identical=len({i["digest"] for i in g}) == 1,
                                     errs=sum((i["errs"] for i in g), []), loss=st["global_loss"])
        torch.cuda.synchronize(); dist.barrier()
        del eng
        torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        print("RESULT " + json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
