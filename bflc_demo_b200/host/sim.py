"""In-process simulator: N clients + sponsor against one C++ ledger, round-robin polling with
no sleeps (the reference needs >= 2 sequential 10-30 s sleeps per round, M:231-233).
``python -m bflc_demo_b200.host.sim`` reproduces the reference demo end to end on the UCI
Occupancy CSV (20 clients, committee 4, top-6 of 10, lr 1e-3, softmax regression)."""
from __future__ import annotations

import argparse
import time
from typing import List, Optional

from .._native import ledger as _ledger
from ..config import FLConfig
from ..data.occupancy import split_data
from ..data.synthetic import Shard, femnist_like
from .client import Client, Sponsor
from .models import HostModel


def build(cfg: FLConfig, shards: List[Shard], test: Optional[Shard], *, model: HostModel,
          genesis=None, log=None, sponsor_log=None):
    L = _ledger()
    led = L.Ledger(cfg.to_ledger_config(model.size))
    if genesis is not None:
        # non-zero genesis model: registered through a zero-lr "round -1" is not possible, so the
        # simulator seeds clients' first global via an offset applied on both sides.
        raise NotImplementedError
    clients = [Client(i, led, shards[i], model, lr=cfg.learning_rate, batch_size=cfg.batch_size,
                      max_epoch=cfg.max_epoch, byzantine=i in cfg.byzantine_ranks,
                      byzantine_scale=cfg.byzantine_scale, log=log) for i in range(cfg.clients)]
    sponsor = Sponsor(led, test, model, log=sponsor_log) if test is not None else None
    return led, clients, sponsor


def run(cfg: FLConfig, shards, test, *, model: HostModel, rounds: int, log=print):
    led, clients, sponsor = build(cfg, shards, test, model=model, log=None, sponsor_log=log)
    t0 = time.time()
    while led.epoch() < rounds:
        progressed = False
        for c in clients:
            if c.poll() not in ("idle", "done"):
                progressed = True
        if sponsor:
            sponsor.poll()
        for line in led.drain_log():
            if log and "global loss" in line:
                log(line)
        if not progressed and led.epoch() >= 0:
            # nobody could act: a stalled round (e.g. dead committee member)
            raise RuntimeError(f"round {led.epoch()} stalled: update_count={led.update_count()} "
                               f"score_count={led.score_count()}")
    dt = time.time() - t0
    return led, clients, sponsor, dt


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=10)
    ap.add_argument("--dataset", default="occupancy", choices=["occupancy", "femnist"])
    ap.add_argument("--clients", type=int, default=20)
    a = ap.parse_args(argv)
    if a.dataset == "occupancy":
        cfg = FLConfig.reference_scaled(a.clients)
        shards, test, src = split_data(clients_num=cfg.clients)
        model = HostModel("softmax", 5, 2)
        print(f"data: {src}; {cfg.clients} clients, committee {cfg.committee_size}, "
              f"top-{cfg.aggregate_count} of {cfg.needed_updates}")
    else:
        cfg = FLConfig.for_world(a.clients, learning_rate=0.05, batch_size=50)
        shards = femnist_like(cfg.clients, 300, seed=1)
        test = femnist_like(1, 1000, seed=1, only=0)[0]
        model = HostModel("mlp", 784, 62, hidden=64, scale_inputs=1 / 255.0)
    led, clients, sponsor, dt = run(cfg, shards, test, model=model, rounds=a.rounds)
    print(f"{a.rounds} rounds in {dt:.2f} s ({a.rounds / dt:.1f} rounds/s); chain ok="
          f"{led.verify_chain()} blocks={led.n_blocks()} counters={led.counters()}")


if __name__ == "__main__":
    main()
