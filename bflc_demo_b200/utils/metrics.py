"""Structured metrics / logging.  The reference prints: contract-side ``std::clog`` lines under
``#if OUTPUT`` (H:4; C:240-242,255-257,291-293,422-425) and client-side ``print`` (M:97,104,210,
241,327-328); its two metrics are the global loss (mean avg_cost of the aggregated trainers) and
the sponsor's test accuracy; gas op counters are the only structured counters.
``RunLog`` keeps the same two metrics per epoch plus ledger counters, as JSON lines."""
from __future__ import annotations

import json
import sys
import time
from typing import Optional, TextIO


class RunLog:
    def __init__(self, stream: Optional[TextIO] = None, path: Optional[str] = None, rank: int = 0):
        self.rank = rank
        self.stream = stream if stream is not None else sys.stdout
        self.file = open(path, "a") if path else None
        self.t0 = time.time()
        self.rows = []

    def round(self, epoch: int, global_loss: float, test_acc: Optional[float] = None, **extra):
        row = dict(t=round(time.time() - self.t0, 4), rank=self.rank, epoch=epoch,
                   global_loss=float(global_loss), **extra)
        if test_acc is not None:
            row["test_acc"] = float(test_acc)
        self.rows.append(row)
        line = json.dumps(row)
        if self.file:
            self.file.write(line + "\n")
            self.file.flush()
        if self.rank == 0 and self.stream:
            # the reference's two human-readable lines (C:424, M:327)
            print("the %d epoch , global loss : %.6f" % (epoch, global_loss), file=self.stream)
            if test_acc is not None:
                print("Epoch: %03d, test_acc: %.4f" % (epoch, test_acc), file=self.stream)
        return row

    def counters(self, ledger) -> dict:
        c = dict(ledger.counters())
        c.update(blocks=ledger.n_blocks(), epoch=ledger.epoch())
        return c

    def close(self):
        if self.file:
            self.file.close()
