"""Tracing / profiling hooks.  The reference has none (one TRACE log of the raw call parameter,
CommitteePrecompiled.cpp:136-137).  Here:

* ``PhaseTimer``   device-side timing of named phases with CUDA events (the headline metric is
                   device-timed, max over ranks), summarised per phase;
* ``nvtx_range``   NVTX ranges so ``ncu``/``nsys`` captures are attributable to protocol phases;
* ``ChromeTrace``  host-side spans dumped as a chrome://tracing JSON file.
"""
from __future__ import annotations

import contextlib
import json
import os
import threading
import time
from collections import defaultdict
from typing import Dict, List, Optional

import torch


@contextlib.contextmanager
def nvtx_range(name: str):
    on = torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class PhaseTimer:
    """``with timer.phase("train"): ...`` records a CUDA-event pair on the current stream.
    ``summary()`` synchronises once and returns {phase: {count, total_ms, mean_ms, max_ms}}."""

    def __init__(self, enabled: bool = True):
        self.enabled = enabled and torch.cuda.is_available()
        self._pairs: Dict[str, List[tuple]] = defaultdict(list)

    @contextlib.contextmanager
    def phase(self, name: str):
        if not self.enabled:
            yield
            return
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        with nvtx_range(name):
            yield
        e1.record()
        self._pairs[name].append((e0, e1))

    def summary(self, reset: bool = True) -> Dict[str, dict]:
        if self.enabled:
            torch.cuda.synchronize()
        out = {}
        for name, pairs in self._pairs.items():
            ms = [a.elapsed_time(b) for a, b in pairs]
            out[name] = dict(count=len(ms), total_ms=sum(ms), mean_ms=sum(ms) / len(ms), max_ms=max(ms))
        if reset:
            self._pairs.clear()
        return out


class ChromeTrace:
    def __init__(self, path: Optional[str] = None, rank: int = 0):
        self.path, self.rank = path, rank
        self.events: List[dict] = []
        self._lock = threading.Lock()

    @contextlib.contextmanager
    def span(self, name: str, **args):
        t0 = time.perf_counter_ns()
        try:
            yield
        finally:
            t1 = time.perf_counter_ns()
            with self._lock:
                self.events.append(dict(name=name, ph="X", ts=t0 / 1e3, dur=(t1 - t0) / 1e3,
                                        pid=self.rank, tid=threading.get_ident() % 100000, args=args))

    def dump(self, path: Optional[str] = None) -> str:
        path = path or self.path or f"trace_rank{self.rank}.json"
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        with open(path, "w") as f:
            json.dump({"traceEvents": self.events}, f)
        return path
