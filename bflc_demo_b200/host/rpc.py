"""Ledger-as-a-service: one process hosts the C++ ledger, clients talk to it over
``multiprocessing.connection`` (authenticated local sockets).  This is the stand-in for the
reference's L2 wire layer -- BcosClient ``call`` / ``sendRawTransactionGetReceipt`` over the
TLS Channel protocol with per-client ECDSA identities (python-sdk/main.py:13-17, 94-96;
README.md:238-260, 348-359): identity = the authenticated connection's client id, payloads
are binary numpy arrays instead of JSON-in-ABI strings."""
from __future__ import annotations

import threading
from multiprocessing.connection import Client as _Conn, Listener
from typing import Optional, Tuple

import numpy as np

from .._native import ledger as _ledger
from ..config import FLConfig

AUTH = b"bflc-demo-b200"
METHODS = ("RegisterNode", "QueryState", "QueryGlobalModel", "UploadLocalUpdate", "UploadScores",
           "QueryAllUpdates", "epoch", "n_blocks", "verify_chain", "counters", "state_hash",
           "drain_log", "snapshot", "last_global_loss")


class LedgerServer:
    def __init__(self, cfg: FLConfig, model_size: int, address: Tuple[str, int] = ("127.0.0.1", 0)):
        self.ledger = _ledger().Ledger(cfg.to_ledger_config(model_size))
        self.listener = Listener(address, authkey=AUTH)
        self.address = self.listener.address
        self._stop = threading.Event()
        self._threads = []

    def _serve(self, conn):
        L = _ledger()
        try:
            while not self._stop.is_set():
                try:
                    msg = conn.recv()
                except (EOFError, ConnectionResetError):
                    break
                name, args = msg
                if name == "__shutdown__":
                    self._stop.set()
                    conn.send(("ok", None))
                    break
                if name not in METHODS:
                    conn.send(("err", f"unknown method {name}"))  # C:312-318
                    continue
                try:
                    out = getattr(self.ledger, name)(*args)
                    if isinstance(out, L.Status):
                        out = ("status", int(out), L.status_name(out))
                    conn.send(("ok", out))
                except Exception as e:  # noqa: BLE001
                    conn.send(("err", repr(e)))
        finally:
            conn.close()

    def serve_forever(self):
        self.listener._listener._socket.settimeout(0.2)
        while not self._stop.is_set():
            try:
                conn = self.listener.accept()
            except Exception:  # timeout
                continue
            t = threading.Thread(target=self._serve, args=(conn,), daemon=True)
            t.start()
            self._threads.append(t)
        self.listener.close()


class _Status(int):
    name = "OK"

    def __new__(cls, v, name):
        o = int.__new__(cls, v)
        o.name = name
        return o


class RemoteLedger:
    """Client-side proxy with the six contract methods."""

    def __init__(self, address):
        self.conn = _Conn(tuple(address), authkey=AUTH)
        self.lock = threading.Lock()

    def _call(self, name, *args):
        with self.lock:
            self.conn.send((name, args))
            kind, out = self.conn.recv()
        if kind == "err":
            raise RuntimeError(out)
        if isinstance(out, tuple) and len(out) == 3 and out[0] == "status":
            return _Status(out[1], out[2])
        return out

    def __getattr__(self, name):
        if name in METHODS:
            return lambda *a: self._call(name, *a)
        raise AttributeError(name)

    def UploadLocalUpdate(self, c, delta, n_samples, avg_cost, ep):
        return self._call("UploadLocalUpdate", c, np.asarray(delta, np.float32), int(n_samples),
                          float(avg_cost), int(ep))

    def shutdown(self):
        try:
            self._call("__shutdown__")
        except Exception:  # noqa: BLE001
            pass

    def finish(self):  # BcosClient.finish()
        self.conn.close()
