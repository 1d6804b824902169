"""Rounds/s of the other BASELINE.json configs (LeNet-5 fp8, ResNet-18 + Byzantine rank,
BERT-base) on the generic engine.  One JSON line per config (rank 0).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node=N --master-addr 127.0.0.1 \
      scripts/bench_models.py [--configs mlp_fp8 lenet5_fp8 resnet18_byz bert] [--rounds 6]

Timing: 2 warm-up rounds, then K rounds bracketed by barrier + synchronize, CUDA events on the
default stream, max over ranks.  (The flagship number is bench.py; this covers model families.)
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from bflc_demo_b200.config import FLConfig
from bflc_demo_b200.data.synthetic import cifar_like, femnist_like, tokens_like
from bflc_demo_b200.engine.generic import GenericFedEngine
from bflc_demo_b200.models.nets import build_model

CONFIGS = {
    #  name          model       dtype   samples batch  lr     bert_layers byzantine
    "mlp_fp8":      ("mlp",      "fp8",  4096,  512,   0.05,  0,  False),
    "mlp_bf16":     ("mlp",      "bf16", 4096,  512,   0.05,  0,  False),
    "lenet5_fp8":   ("lenet5",   "fp8",  2048,  128,   0.05,  0,  False),
    "lenet5_bf16":  ("lenet5",   "bf16", 2048,  128,   0.05,  0,  False),
    "resnet18_byz": ("resnet18", "bf16", 256,   64,    0.02,  0,  True),
    "bert":         ("bert",     "bf16", 32,    16,    0.002, 12, False),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", nargs="*", default=["mlp_fp8", "lenet5_fp8", "resnet18_byz", "bert"])
    ap.add_argument("--rounds", type=int, default=6)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    lr_ = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr_)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr_))

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for name in a.configs:
        model, dtype, S, B, lr, layers, byz = CONFIGS[name]
        cfg = FLConfig.for_world(world, model=model, batch_size=B, samples_per_client=S,
                                 learning_rate=lr, dtype=dtype, ring_slots=256,
                                 byzantine_ranks=[world - 1] if (byz and world > 2) else [])
        if model == "mlp":
            shard = femnist_like(world, S, seed=7, only=rank)[0]
        elif model in ("lenet5", "resnet18"):
            shard = cifar_like(world, S, seed=7, alpha=0.5)[rank]
        else:
            shard = tokens_like(world, S, seed=7)[rank]
        net = build_model(model, shard.n_classes, layers=layers or 12)
        eng = GenericFedEngine(cfg, net, shard, rank=rank, world=world, device=lr_)
        for _ in range(2):
            eng.run_round()
        sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.rounds):
            eng.run_round()
        e1.record()
        sync()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        st = eng.read_state()
        errs = eng.drain_blocks()
        if rank == 0:
            print(json.dumps({
                "config": name, "model": model, "dtype": dtype, "n_gpus": world,
                "params": int(net.spec.total), "samples_per_client": S, "local_batch": B,
                "committee": cfg.committee_size, "trainers": cfg.n_trainers, "byzantine": cfg.byzantine_ranks,
                "rounds": a.rounds, "ms_per_round": ms.item() / a.rounds,
                "rounds_per_s": a.rounds / (ms.item() / 1e3), "global_loss": st["global_loss"],
                "two_shot": bool(eng.two_shot), "multicast": eng.heap.describe().get("multicast"),
                "ledger_mismatches": errs[:2], "chain_ok": eng.host_ledger.verify_chain()}), flush=True)
        del eng
        torch.cuda.empty_cache()
        sync()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
