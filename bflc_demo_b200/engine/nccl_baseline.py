"""OUR NCCL + cuBLAS baseline of the reference algorithm (NOT a reference-published build:
iammcy/BFLC-demo ships no GPU code and cannot be built offline -- SURVEY.md 0.3).

Same protocol, same model, same data and precision as ``FusedEngine``, but built the
conventional way: PyTorch ops (cuBLAS GEMMs) for local training and validation, NCCL
collectives for the three exchanges, the election on the host:

    all_gather(trainer weights) -> committee validates every candidate -> all_gather(score
    rows) -> median / top-K / sample-weighted FedAvg -> (optional) broadcast of the result

It is tuned the way a competent user would: the local-training pass and the validation pass
are each captured in a CUDA graph, collectives use pre-allocated flat buffers, the
aggregation runs redundantly on every rank so no broadcast is needed unless
``broadcast=True`` (the literal BASELINE.json loop).  What it cannot avoid is the thing the
fused engine is built to remove: host-launched collectives and one device->host read per
round to learn the new roles.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from ..config import FLConfig
from ..data.synthetic import Shard
from ..models.mlp import mlp_spec
from ..protocol import oracle as O
from .fused import ROLE_COMM, ROLE_TRAINER, initial_roles


class NcclBaselineEngine:
    def __init__(self, cfg: FLConfig, shard: Shard, *, rank: int = 0, world: int = 1,
                 device: int = 0, group=None, broadcast: bool = False):
        self.cfg, self.rank, self.world, self.group = cfg, rank, world, group
        self.dev = torch.device("cuda", device)
        torch.cuda.set_device(device)
        self.broadcast = broadcast
        x0 = shard.x.reshape(len(shard), -1)
        self.in_dim = x0.shape[1]
        self.spec = mlp_spec(self.in_dim, cfg.hidden, shard.n_classes)
        P = self.spec.total
        self.P = P
        B = cfg.batch_size
        self.S = (len(shard) // B) * B
        self.steps = (self.S // B) * cfg.local_epochs
        self.n_val = min(cfg.val_samples or len(shard), len(shard))
        init = torch.empty(P)
        self.spec.init_(init, seed=cfg.seed + 1234)
        self.global_w = init.to(self.dev)
        self.work = self.global_w.clone()
        self.wv = self.spec.views(self.work)
        self.all_w = torch.empty(world, P, device=self.dev)
        self.all_scores = torch.zeros(world, world, device=self.dev)
        self.my_scores = torch.zeros(world, device=self.dev)
        self.all_meta = torch.zeros(world, 2, device=self.dev)
        self.my_meta = torch.zeros(2, device=self.dev)
        self.roles: List[int] = initial_roles(cfg)
        self.epoch = 0
        self.global_loss = 0.0
        self.host_x = x0.contiguous().pin_memory()
        self.host_y = shard.y.to(torch.int64).contiguous().pin_memory()
        self.x_u8 = self.host_x.to(self.dev)
        self.y = self.host_y.to(self.dev)
        self.x_bf = torch.empty(len(shard), self.in_dim, device=self.dev, dtype=torch.bfloat16)
        self.loss_acc = torch.zeros(1, device=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.train_graph: Optional[torch.cuda.CUDAGraph] = None
        self.val_graph: Optional[torch.cuda.CUDAGraph] = None
        self.out_host = torch.zeros(world + 2, dtype=torch.float32).pin_memory()
        self.byz = rank in cfg.byzantine_ranks

    # ---------------------------------------------------------------- local work
    def _train_pass(self):
        cfg = self.cfg
        B, lr = cfg.batch_size, cfg.learning_rate
        w = self.wv
        self.x_bf.copy_(self.x_u8.to(torch.bfloat16) * (1.0 / 255.0))
        self.work.copy_(self.global_w)
        self.loss_acc.zero_()
        for i in range(self.steps):
            x = self.x_bf[i * B:(i + 1) * B]
            y = self.y[i * B:(i + 1) * B]
            w1b, w2b = w["w1"].to(torch.bfloat16), w["w2"].to(torch.bfloat16)
            h = torch.relu(torch.addmm(w["b1"].to(torch.bfloat16), x, w1b.t()))
            logits = torch.addmm(w["b2"].to(torch.bfloat16), h, w2b.t()).float()
            lse = torch.logsumexp(logits, 1)
            self.loss_acc += (lse - logits.gather(1, y[:, None]).squeeze(1)).sum()
            p = torch.softmax(logits, 1)
            p.scatter_add_(1, y[:, None], torch.full((B, 1), -1.0, device=self.dev))
            dl = (p / B).to(torch.bfloat16)
            gw2 = (dl.t() @ h).float()
            gb2 = dl.float().sum(0)
            dh = (dl @ w2b) * (h > 0)
            gw1 = (dh.t() @ x).float()
            gb1 = dh.float().sum(0)
            torch._foreach_add_([w["w1"], w["b1"], w["w2"], w["b2"]], [gw1, gb1, gw2, gb2], alpha=-lr)
        if self.byz:
            self.work.copy_(self.global_w - self.cfg.byzantine_scale * (self.work - self.global_w))
        self.my_meta[0:1].fill_(float(self.S))
        self.my_meta[1:2] = self.loss_acc / float(self.steps * B)

    def _val_pass(self):
        xv = self.x_bf[: self.n_val]
        yv = self.y[: self.n_val]
        W = self.all_w
        e = self.spec.by_name
        def part(name):
            en = e[name]
            return W[:, en.offset:en.offset + en.numel].reshape(self.world, *en.shape)
        w1, b1, w2, b2 = part("w1"), part("b1"), part("w2"), part("b2")
        h = torch.relu(torch.baddbmm(b1.to(torch.bfloat16)[:, None, :],
                                     xv[None].expand(self.world, -1, -1),
                                     w1.to(torch.bfloat16).transpose(1, 2)))
        logits = torch.baddbmm(b2.to(torch.bfloat16)[:, None, :], h,
                               w2.to(torch.bfloat16).transpose(1, 2)).float()
        self.my_scores.copy_((logits.argmax(2) == yv[None]).float().mean(1))

    def capture(self):
        with torch.cuda.stream(self.stream):
            self.x_bf.copy_(self.x_u8.to(torch.bfloat16) * (1.0 / 255.0))
            self.all_w.copy_(self.global_w[None].expand(self.world, -1))
            for _ in range(2):
                self._train_pass()
                self._val_pass()
        self.stream.synchronize()
        if self.cfg.cuda_graph:
            self.train_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.train_graph, stream=self.stream):
                self._train_pass()
            self.val_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.val_graph, stream=self.stream):
                self._val_pass()
        self.work.copy_(self.global_w)
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- one round
    def run_round(self) -> dict:
        cfg, n = self.cfg, self.world
        role = self.roles[self.rank]
        with torch.cuda.stream(self.stream):
            if role & ROLE_TRAINER:
                self.train_graph.replay() if self.train_graph else self._train_pass()
            if n > 1:
                dist.all_gather_into_tensor(self.all_w.view(-1), self.work, group=self.group)
                dist.all_gather_into_tensor(self.all_meta.view(-1), self.my_meta, group=self.group)
            else:
                self.all_w[0].copy_(self.work)
                self.all_meta[0].copy_(self.my_meta)
            if role & ROLE_COMM:
                self.val_graph.replay() if self.val_graph else self._val_pass()
            if n > 1:
                dist.all_gather_into_tensor(self.all_scores.view(-1), self.my_scores, group=self.group)
            else:
                self.all_scores[0].copy_(self.my_scores)
            # election on the host (one small D2H per round)
            scores_h = self.all_scores.cpu()
            meta_h = self.all_meta.cpu()
        trainers = [r for r in range(n) if self.roles[r] & ROLE_TRAINER]
        comm = [r for r in range(n) if self.roles[r] & ROLE_COMM]
        res = O.run_consensus(n, cfg.committee_size, cfg.aggregate_count,
                              {r: self.roles[r] for r in range(n)}, trainers,
                              {c: {t: float(scores_h[c, t]) for t in trainers} for c in comm},
                              {t: int(meta_h[t, 0]) for t in trainers},
                              {t: float(meta_h[t, 1]) for t in trainers}, cfg.weight_by_score)
        with torch.cuda.stream(self.stream):
            if self.broadcast and n > 1:
                if self.rank == 0:
                    self._apply(res)
                dist.broadcast(self.global_w, src=0, group=self.group)
            else:
                self._apply(res)
        self.roles = [res.role_after[r] for r in range(n)]
        self.global_loss = res.global_loss
        self.epoch += 1
        return dict(epoch=self.epoch, roles=list(self.roles), global_loss=res.global_loss,
                    selected=res.selected)

    def _apply(self, res):
        if not res.selected:
            return
        w = torch.tensor([res.weight[t] for t in res.selected], device=self.dev)
        idx = torch.tensor(res.selected, device=self.dev)
        self.global_w.copy_((self.all_w.index_select(0, idx) * w[:, None]).sum(0))

    def run_round_e2e(self, host_x=None, host_y=None) -> dict:
        hx = self.host_x if host_x is None else host_x
        hy = self.host_y if host_y is None else host_y
        with torch.cuda.stream(self.stream):
            self.x_u8.copy_(hx, non_blocking=True)
            self.y.copy_(hy, non_blocking=True)
        out = self.run_round()
        self.stream.synchronize()
        return out

    @property
    def h2d_bytes_per_round(self) -> int:
        return self.host_x.numel() + self.host_y.numel() * 8

    @property
    def d2h_bytes_per_round(self) -> int:
        return (self.world * self.world + self.world * 2) * 4
