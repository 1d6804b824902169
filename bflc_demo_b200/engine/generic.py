"""Model-agnostic federated engine: any ``FlatNet`` (LeNet-5, ResNet-18, BERT-base, MLP) on the
same device-resident protocol as ``FusedEngine`` -- symmetric-heap upload buffers, epoch-tagged
P2P flags, the consensus/aggregation kernel, the host ledger re-executing every election --
with the round driven from Python instead of a captured graph (rounds of these models take
milliseconds, so one device->host read of the role table per round is noise).

Committee validation runs the model *directly on the trainers' HBM*: a candidate's ``Bound`` is a
set of tensor views over the peer-mapped upload buffers, so every GEMM of the forward pass
TMA-loads its weight tiles across NVLink -- the QueryAllUpdates all-gather (reference
C:299-311, M:196-217) never materialises.
"""
from __future__ import annotations

import struct
from typing import List, Optional

import torch
import torch.distributed as dist

from .._native import C, ledger as _ledger
from ..config import FLConfig
from ..data.synthetic import Shard
from ..models.nets import Bound, FlatNet
from ..parallel.layout import HeapLayout
from ..parallel.symm import SymmetricHeap
from .fused import ROLE_COMM, ROLE_TRAINER, FusedEngine, initial_roles


class GenericFedEngine:
    read_state = FusedEngine.read_state
    drain_blocks = FusedEngine.drain_blocks

    def __init__(self, cfg: FLConfig, net: FlatNet, shard: Shard, *, rank: int = 0, world: int = 1,
                 device: int = 0, group=None):
        assert cfg.clients == world and world <= 8
        self.cfg, self.net, self.rank, self.world, self.device = cfg, net, rank, world, device
        self.group = group
        torch.cuda.set_device(device)
        self.dev = torch.device("cuda", device)
        self.mod = C()
        # cfg.dtype "fp8": forward GEMMs of Linear / Conv2d run block-scaled fp8 (ops/mx8.py)
        from ..ops import nn as _nn
        _nn.set_precision("mx8" if cfg.dtype == "fp8" else "bf16")
        self.sz = sz = self.mod.struct_sizes()
        self.spec = net.spec
        P = self.n_params = net.spec.total
        B = cfg.batch_size
        self.S = (len(shard) // B) * B
        self.steps = (self.S // B) * cfg.local_epochs
        self.n_val = min(cfg.val_samples or len(shard), len(shard))
        self.layout = HeapLayout(P, cfg.ring_slots)
        self.heap = SymmetricHeap(self.layout.total_bytes, rank=rank, world=world, device=device,
                                  group=group, want_multicast=cfg.use_multicast)
        self.fed = self.layout.fed_dict(rank, world, self.heap.peer_ptrs, self.heap.mc_ptr)
        o, hv = self.layout.offsets, self.heap.view
        self.work_master = hv(o["work_master"], [P], torch.float32)
        self.work_shadow = hv(o["work_shadow"], [P], torch.bfloat16)
        self.global_master = hv(o["global"], [P], torch.float32)
        self.global_shadow = hv(o["global_shadow"], [P], torch.bfloat16)
        self.state_bytes = hv(o["state"], [sz["RoundState"]], torch.uint8)
        self.ring_bytes = hv(o["ring"], [cfg.ring_slots * sz["BlockRecord"]], torch.uint8)
        self.loss_sum = hv(o["plan"] + sz["plan_loss_sum_off"], [1], torch.float32)
        self.val_correct = hv(o["plan"] + sz["plan_correct_off"], [sz["kMaxRanks"]], torch.int32)
        self.opt_step_ptr = self.heap.local_ptr + o["plan"] + sz["plan_opt_step_off"]
        self.grad = torch.zeros(P, device=self.dev)
        self.m = torch.zeros(P, device=self.dev) if cfg.optimizer == "adam" else None
        self.v = torch.zeros(P, device=self.dev) if cfg.optimizer == "adam" else None

        init = torch.empty(P)
        net.init_(init, seed=cfg.seed + 1234)
        for t in (self.work_master, self.global_master):
            t.copy_(init)
        for t in (self.work_shadow, self.global_shadow):
            t.copy_(init.to(torch.bfloat16))
        self.bound = net.bind(self.work_master, self.work_shadow, self.grad)

        roles = initial_roles(cfg)
        st = self.mod.state_init_bytes(world, cfg.committee_size, cfg.aggregate_count, roles)
        self.state_bytes.copy_(torch.frombuffer(bytearray(st), dtype=torch.uint8))
        self.host_ledger = _ledger().Ledger(cfg.to_ledger_config(P))
        self.host_ledger.Bootstrap(roles)
        self.drained = 0

        self.x = net.preprocess(shard.x.to(self.dev))
        self.y = shard.y.to(self.dev, torch.int32)
        self.two_shot = cfg.two_shot if cfg.two_shot is not None else (P * 4 > (64 << 20) and world > 1)
        self.byz = 1 if rank in cfg.byzantine_ranks else 0
        self._peer_bounds = {}
        self._stage = None
        if world > 1:
            dist.barrier(group=group)
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ pieces
    def peer_bound(self, t: int, parity: int) -> Bound:
        key = (t, parity)
        if key not in self._peer_bounds:
            o, P = self.layout.offsets, self.n_params
            master = self.heap.view(o[f"upload_master{parity}"], [P], torch.float32, rank=t)
            shadow = self.heap.view(o[f"upload_shadow{parity}"], [P], torch.bfloat16, rank=t)
            self._peer_bounds[key] = self.net.bind(master, shadow, None)
        return self._peer_bounds[key]

    def local_training(self):
        cfg, B = self.cfg, self.cfg.batch_size
        for i in range(self.steps):
            j = (i * B) % self.S
            loss = self.net.loss(self.bound, self.x[j:j + B], self.y[j:j + B])
            loss.backward()
            self.loss_sum += loss.detach() * B
            self.mod.optim_step(cfg.optimizer == "adam", self.work_master, self.grad,
                                self.work_shadow, self.m, self.v, cfg.learning_rate, 0.0, 0.9,
                                0.999, 1e-8, i + 1, self.opt_step_ptr, 0, True)

    def validate(self, trainers: List[int], parity: int):
        xv, yv = self.x[: self.n_val], self.y[: self.n_val]
        if self.cfg.stage_candidates and self.world > 1:
            # one P2P pass per candidate (weights + fp32 master) into local staging, started per
            # candidate as soon as its trainer's flag is up; the forward passes then read local HBM
            if self._stage is None:
                P = self.n_params
                self._stage = (torch.empty(self.world, P, device=self.dev, dtype=torch.bfloat16),
                               torch.empty(self.world, P, device=self.dev, dtype=torch.float32))
                self._stage_bounds = [self.net.bind(self._stage[1][z], self._stage[0][z], None)
                                      for z in range(self.world)]
            self.mod.fed_pull_candidates(self.fed, self._stage[0], self._stage[1])
            bounds = [self._stage_bounds[z] for z in range(len(trainers))]
        else:
            # direct: every GEMM of the forward pass TMA-loads its weight tiles from the peer
            self.mod.fed_wait_trained(self.fed)
            bounds = [self.peer_bound(t, parity) for t in trainers]
        for z, b in enumerate(bounds):
            cnt = self.net.correct(b, xv, yv)
            self.val_correct[z:z + 1].copy_(cnt)

    # ------------------------------------------------------------------ one round
    def run_round(self) -> dict:
        m, cfg = self.mod, self.cfg
        # ring backpressure: drain the device BlockRecord ring before slots can be overwritten
        self._rounds = getattr(self, "_rounds", 0) + 1
        if self._rounds - self.drained >= max(cfg.ring_slots // 2, 1):
            errs = self.drain_blocks()
            if errs:
                raise RuntimeError(f"host/device ledgers disagree: {errs[:2]}")
        m.fed_plan_round(self.fed, [], self.steps, False)
        st = self.read_state()  # host learns roles/epoch (D2H of the 104-byte ledger page)
        role = st["roles"][self.rank]
        trainers = [r for r in range(self.world) if st["roles"][r] & ROLE_TRAINER]
        if role & ROLE_TRAINER:
            self.local_training()
        m.fed_upload(self.fed, self.S, self.steps * cfg.batch_size, self.byz, cfg.byzantine_scale)
        if role & ROLE_COMM:
            self.validate(trainers, st["epoch"] & 1)
        m.fed_consensus_aggregate(self.fed, self.n_val, cfg.weight_by_score, self.two_shot,
                                  cfg.use_multicast and self.heap.has_multicast)
        return st

    def evaluate(self, shard: Shard) -> float:
        x = self.net.preprocess(shard.x.to(self.dev))
        y = shard.y.to(self.dev, torch.int32)
        b = self.net.bind(self.global_master, self.global_shadow, None)
        return float(self.net.correct(b, x, y).item()) / len(shard)
