"""State-machine replication of the ledger over ``torch.distributed`` (gloo on CPU, also
works on nccl): every rank keeps a full C++ ledger replica; state-changing calls are queued
and, once per tick, all-gathered and applied by every replica in the same (tick, rank, seq)
order -- the role PBFT plays for the reference (4 chain nodes re-executing every tx,
README.md:162-168).  After every tick the replicas' state hashes are compared.

This is BASELINE.json config #1 ("4 CPU/gloo clients, committee_size=2, plumbing, no GPU")."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch.distributed as dist

from .._native import ledger as _ledger
from ..config import FLConfig
from ..data.synthetic import Shard
from .client import Client, Sponsor
from .models import HostModel


class ReplicatedLedger:
    def __init__(self, cfg: FLConfig, model_size: int, group=None, check_every: int = 1):
        self.L = _ledger()
        self.replica = self.L.Ledger(cfg.to_ledger_config(model_size))
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.queue: List[tuple] = []
        self.tick = 0
        self.check_every = check_every
        self.applied = 0

    # ---- views: served by the local replica (reference: client.call, M:106,198,207,245)
    def QueryState(self, c):
        return self.replica.QueryState(c)

    def QueryGlobalModel(self):
        return self.replica.QueryGlobalModel()

    def QueryAllUpdates(self):
        return self.replica.QueryAllUpdates()

    def epoch(self):
        return self.replica.epoch()

    # ---- transactions: queued, ordered and applied by sync() (sendRawTransactionGetReceipt)
    def RegisterNode(self, c):
        self.queue.append(("reg", c))
        return self.L.Status.OK

    def UploadLocalUpdate(self, c, delta, n_samples, avg_cost, ep):
        self.queue.append(("upd", c, np.asarray(delta, np.float32), int(n_samples), float(avg_cost), int(ep)))
        return self.L.Status.OK

    def UploadScores(self, c, ep, scores):
        self.queue.append(("sco", c, int(ep), {int(k): float(v) for k, v in scores.items()}))
        return self.L.Status.OK

    def sync(self) -> List:
        mine, self.queue = self.queue, []
        if self.world > 1:
            allq: List[Optional[list]] = [None] * self.world
            dist.all_gather_object(allq, mine, group=self.group)
        else:
            allq = [mine]
        results = []
        for r, txs in enumerate(allq):        # deterministic total order: rank-major per tick
            for tx in txs:
                if tx[0] == "reg":
                    st = self.replica.RegisterNode(tx[1])
                elif tx[0] == "upd":
                    st = self.replica.UploadLocalUpdate(tx[1], tx[2], tx[3], tx[4], tx[5])
                else:
                    st = self.replica.UploadScores(tx[1], tx[2], tx[3])
                self.applied += 1
                if r == self.rank:
                    results.append(st)
        self.tick += 1
        if self.world > 1 and self.check_every and self.tick % self.check_every == 0:
            hs: List[Optional[str]] = [None] * self.world
            dist.all_gather_object(hs, self.replica.state_hash(), group=self.group)
            if len(set(hs)) != 1:
                raise RuntimeError(f"ledger replicas diverged at tick {self.tick}: {hs}")
        return results


def run_replicated(cfg: FLConfig, shard: Shard, test: Optional[Shard], model: HostModel,
                   rounds: int, group=None, log=None):
    """One client per rank; returns (ledger, client, sponsor history)."""
    led = ReplicatedLedger(cfg, model.size, group)
    me = Client(led.rank, led, shard, model, lr=cfg.learning_rate, batch_size=cfg.batch_size,
                byzantine=led.rank in cfg.byzantine_ranks, byzantine_scale=cfg.byzantine_scale)
    sponsor = Sponsor(led, test, model, log=log) if (test is not None and led.rank == 0) else None
    guard = 0
    while led.epoch() < rounds:
        me.poll()
        led.sync()
        if sponsor:
            sponsor.poll()
        guard += 1
        if guard > 50 * (rounds + 2):
            raise RuntimeError("replicated run made no progress")
    return led, me, sponsor
