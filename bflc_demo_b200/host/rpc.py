"""Ledger-as-a-service: one process hosts the C++ ledger, clients talk to it over
``multiprocessing.connection`` sockets.  This is the stand-in for the reference's L2 wire layer
-- BcosClient ``call`` / ``sendRawTransactionGetReceipt`` over the TLS Channel protocol with
per-client ECDSA identities (python-sdk/main.py:13-17, 94-96; README.md:238-260, 348-359).

Caller identity (reference: the signed transaction's origin address, CommitteePrecompiled.cpp:147)
is bound PER CONNECTION by a handshake and injected by the server; a client can never name the
id it acts for in a request:

  * signed mode (``accounts`` given -- what ``host/launch.py`` uses): the server sends a random
    nonce, the client answers with its account address and an ECDSA signature over the nonce made
    with its ``node_<i>.pem`` key (``host/identity.py``); the server looks the address up in its
    ``AccountRegistry``, verifies the signature and pins the enrolled client id to the connection.
  * open mode (no accounts; tests and single-user runs): the client claims an id once in the
    handshake and the connection is pinned to it -- still no per-request id.

The socket itself is protected by a per-launch random ``authkey`` (HMAC challenge of
``multiprocessing.connection``, completed before anything is unpickled); the launcher hands it to
its children.  Payloads are binary numpy arrays instead of JSON-in-ABI strings."""
from __future__ import annotations

import os
import threading
from multiprocessing.connection import Client as _Conn, Listener
from typing import Dict, Optional, Tuple

import numpy as np

from .._native import ledger as _ledger
from ..config import FLConfig
from . import identity as _identity

METHODS = ("RegisterNode", "QueryState", "QueryGlobalModel", "UploadLocalUpdate", "UploadScores",
           "QueryAllUpdates", "epoch", "n_blocks", "verify_chain", "counters", "state_hash",
           "drain_log", "snapshot", "last_global_loss")
# methods whose first ledger argument is the caller's client id: supplied by the server
CALLER_METHODS = frozenset(("RegisterNode", "QueryState", "UploadLocalUpdate", "UploadScores"))
OBSERVER = -1   # connection id of a read-only caller (the sponsor)


class LedgerServer:
    def __init__(self, cfg: FLConfig, model_size: int, address: Tuple[str, int] = ("127.0.0.1", 0),
                 *, accounts: Optional[Dict[int, object]] = None, authkey: Optional[bytes] = None):
        """``accounts``: {client id: ECDSA public key} -> signed mode.  ``authkey``: socket secret
        (default: 32 fresh random bytes, readable as ``self.authkey``)."""
        self.cfg = cfg
        self.ledger = _ledger().Ledger(cfg.to_ledger_config(model_size))
        self.authkey = authkey if authkey is not None else os.urandom(32)
        self.listener = Listener(address, authkey=self.authkey)
        self.address = self.listener.address
        self.registry: Optional[_identity.AccountRegistry] = None
        if accounts is not None:
            self.registry = _identity.AccountRegistry()
            for cid, pub in accounts.items():
                self.registry.enroll(int(cid), pub)
        self._stop = threading.Event()
        self._threads = []

    # -- handshake: returns the client id pinned to this connection, or raises PermissionError
    def _handshake(self, conn) -> int:
        nonce = os.urandom(32)
        conn.send(("nonce", nonce))
        msg = conn.recv()
        if not (isinstance(msg, tuple) and len(msg) == 3 and msg[0] == "__hello__"):
            raise PermissionError("handshake expected")
        _, who, sig = msg
        if self.registry is not None:
            if who is None:                      # unsigned observer: views only
                return OBSERVER
            return self.registry.authenticate(str(who), nonce, bytes(sig))
        if who is None:
            return OBSERVER
        cid = int(who)
        if not (0 <= cid < self.cfg.clients):
            raise PermissionError(f"client id {cid} out of range")
        return cid

    def _serve(self, conn):
        L = _ledger()
        try:
            try:
                cid = self._handshake(conn)
                conn.send(("ok", cid))
            except (PermissionError, EOFError, ConnectionResetError, ValueError) as e:
                try:
                    conn.send(("err", f"handshake rejected: {e}"))
                except Exception:  # noqa: BLE001
                    pass
                return
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
                    if name in CALLER_METHODS:
                        if cid == OBSERVER:
                            raise PermissionError(f"{name} needs a client identity")
                        args = (cid,) + tuple(args)      # identity = the connection, C:147
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
            except Exception:  # timeout / failed authkey challenge
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
    """Client-side proxy with the six contract methods.  ``client_id`` / ``key`` identify the
    caller ONCE, at connect time: ``key`` (an ECDSA private key, ``identity.load_account``) signs
    the server's nonce in signed mode; with neither the connection is a read-only observer.

    The id-taking methods keep the in-process ledger's signatures (``RegisterNode(c)``, ...) so
    ``host/client.py`` runs unchanged against either; the id argument must equal the pinned one
    and is NOT transmitted."""

    def __init__(self, address, *, authkey: bytes, client_id: Optional[int] = None, key=None):
        self.conn = _Conn(tuple(address), authkey=authkey)
        self.lock = threading.Lock()
        kind, nonce = self.conn.recv()
        assert kind == "nonce"
        if key is not None:
            addr = _identity.address_of(key.public_key())
            self.conn.send(("__hello__", addr, _identity.sign(key, nonce)))
        else:
            self.conn.send(("__hello__", client_id, b""))
        kind, out = self.conn.recv()
        if kind == "err":
            self.conn.close()
            raise PermissionError(out)
        self.client_id = int(out)
        if client_id is not None and self.client_id != int(client_id):
            self.conn.close()
            raise PermissionError(f"server pinned id {self.client_id}, expected {client_id}")

    def _call(self, name, *args):
        with self.lock:
            self.conn.send((name, args))
            kind, out = self.conn.recv()
        if kind == "err":
            raise RuntimeError(out)
        if isinstance(out, tuple) and len(out) == 3 and out[0] == "status":
            return _Status(out[1], out[2])
        return out

    def _mine(self, c):
        if int(c) != self.client_id:
            raise PermissionError(f"connection is pinned to client {self.client_id}, not {c}")

    def __getattr__(self, name):
        if name in METHODS and name not in CALLER_METHODS:
            return lambda *a: self._call(name, *a)
        raise AttributeError(name)

    def RegisterNode(self, c):
        self._mine(c)
        return self._call("RegisterNode")

    def QueryState(self, c):
        self._mine(c)
        return self._call("QueryState")

    def UploadLocalUpdate(self, c, delta, n_samples, avg_cost, ep):
        self._mine(c)
        return self._call("UploadLocalUpdate", np.asarray(delta, np.float32), int(n_samples),
                          float(avg_cost), int(ep))

    def UploadScores(self, c, ep, scores):
        self._mine(c)
        return self._call("UploadScores", int(ep), scores)

    def shutdown(self):
        try:
            self._call("__shutdown__")
        except Exception:  # noqa: BLE001
            pass

    def finish(self):  # BcosClient.finish()
        self.conn.close()
