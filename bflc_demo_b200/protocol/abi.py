"""The reference's client-facing call surface, kept name-for-name.

The reference client talks to the contract through exactly two SDK verbs:
``client.call(addr, abi, fn[, args])`` for views and
``client.sendRawTransactionGetReceipt(addr, abi, fn, args)`` for state changes
(python-sdk/main.py:106,160,198,207,219,240,245,320), with JSON strings as payloads
(``serialize``/``deserialize``, M:22-30).  ``ContractClient`` offers the same two verbs and the
same JSON payload format on top of ANY ledger object of this package (local C++ ledger, gloo
replica, RPC proxy), so code written against the reference's API ports by changing one import.
"""
from __future__ import annotations

import json
from typing import Any, Dict, List, Optional

import numpy as np

from .._native import ledger as _ledger

CONTRACT_ADDRESS = "0x0000000000000000000000000000000000005006"  # M:80, README.md:63-64
ROLE_NAMES = {1: "trainer", 2: "comm", 3: "comm"}               # M:58-59


def serialize(obj: Any) -> str:      # M:23-25
    return json.dumps(obj)


def deserialize(text: str) -> Any:   # M:28-30
    return json.loads(text)


def methods() -> List[tuple]:
    """[(id, solidity signature, is_view)] -- CommitteePrecompiled.sol:3-10."""
    return list(_ledger().method_table())


def pack_update(delta_W, delta_b, n_samples: int, avg_cost: float) -> str:
    """The reference's update JSON (M:153-158): {'delta_model': {'ser_W', 'ser_b'}, 'meta': ...}."""
    return serialize({"delta_model": {"ser_W": np.asarray(delta_W).tolist(),
                                      "ser_b": np.asarray(delta_b).tolist()},
                      "meta": {"n_samples": int(n_samples), "avg_cost": float(avg_cost)}})


def unpack_update(text: str):
    u = deserialize(text)
    dm, meta = u["delta_model"], u["meta"]
    W, b = np.asarray(dm["ser_W"], np.float32), np.asarray(dm["ser_b"], np.float32)
    return W, b, int(meta["n_samples"]), float(meta["avg_cost"])


class ContractClient:
    """``BcosClient`` look-alike bound to one client identity (the tx origin, C:147)."""

    def __init__(self, ledger, node_id: int, n_features: int = 5, n_class: int = 2):
        self.ledger, self.node_id = ledger, node_id
        self.n_features, self.n_class = n_features, n_class

    # the flat weight vector of this package stores W as [out, in]; the reference's JSON is [in][out]
    def _flat(self, W, b):
        return np.concatenate([np.asarray(W, np.float32).T.reshape(-1), np.asarray(b, np.float32)])

    def _model_json(self, flat) -> str:
        flat = np.asarray(flat, np.float32)
        n = self.n_features * self.n_class
        W = flat[:n].reshape(self.n_class, self.n_features).T
        return serialize({"ser_W": W.tolist(), "ser_b": flat[n:n + self.n_class].tolist()})

    def call(self, to_address: str, abi: Any, fn: str, args: Optional[list] = None):
        L = _ledger()
        m = L.method_from_signature(fn)
        if m == 1:      # QueryState -> (role, epoch)
            role, ep = self.ledger.QueryState(self.node_id)
            return ROLE_NAMES.get(int(role), "trainer"), ep
        if m == 2:      # QueryGlobalModel -> (json, epoch)
            w, ep = self.ledger.QueryGlobalModel()
            return self._model_json(w), ep
        if m == 5:      # QueryAllUpdates -> (json map addr -> update-json, or "")
            ups = self.ledger.QueryAllUpdates()
            if len(ups) == 0:
                return ("",)
            out = {}
            n = self.n_features * self.n_class
            for u in ups:
                d = np.asarray(u["delta"], np.float32)
                out[str(u["sender"])] = pack_update(d[:n].reshape(self.n_class, self.n_features).T,
                                                    d[n:n + self.n_class], u["n_samples"], u["avg_cost"])
            return (serialize(out),)
        raise ValueError(f"{fn!r} is not a view method of the contract")  # C:312-318

    def sendRawTransactionGetReceipt(self, to_address: str, abi: Any, fn: str, args: Optional[list] = None):
        L = _ledger()
        m = L.method_from_signature(fn)
        args = args or []
        if m == 0:
            st = self.ledger.RegisterNode(self.node_id)
        elif m == 3:    # UploadLocalUpdate(update_json, epoch)
            W, b, n, c = unpack_update(args[0])
            st = self.ledger.UploadLocalUpdate(self.node_id, self._flat(W, b), n, c, int(args[1]))
        elif m == 4:    # UploadScores(epoch, scores_json)
            scores = {int(k): float(v) for k, v in deserialize(args[1]).items()}
            st = self.ledger.UploadScores(self.node_id, int(args[0]), scores)
        else:
            raise ValueError(f"{fn!r} is not a transaction method of the contract")
        return {"status": getattr(st, "name", str(st)), "output": ""}

    def finish(self):
        f = getattr(self.ledger, "finish", None)
        if f:
            f()
