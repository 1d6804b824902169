"""The reference's data pipeline (python-sdk/main.py:32-49, component P2): read the UCI
Occupancy CSV, split 75/25 with a fixed seed, one-hot labels, contiguous IID shards via
``np.array_split``.  sklearn/pandas are not required: the split is a seeded permutation
(documented deviation: not bit-identical to ``train_test_split(random_state=42)``, same
proportions).

The table ships in-tree as ``occupancy_uci.npz`` (the public UCI Occupancy Detection training
set, artefact A3: 8143 rows x the five features the reference uses + the binary label, stored
as float32/int8 arrays), so nothing depends on a checkout of the reference.  A CSV in the
reference's format can still be supplied (``path`` / ``BFLC_OCCUPANCY_CSV``); a
schema-compatible synthetic table is the last resort."""
from __future__ import annotations

import csv
import os
from typing import List, Optional, Tuple

import numpy as np
import torch

from .synthetic import Shard, occupancy_like

FEATURES = ["Temperature", "Humidity", "Light", "CO2", "HumidityRatio"]  # M:35-36
IN_TREE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "occupancy_uci.npz")


def load_table(path: Optional[str] = None) -> Tuple[np.ndarray, np.ndarray, str]:
    path = path or os.environ.get("BFLC_OCCUPANCY_CSV", "")
    if path and os.path.exists(path):
        xs, ys = [], []
        with open(path, newline="") as f:
            rd = csv.reader(f)
            header = [name.strip('"') for name in next(rd)]
            first = next(rd)
            # the UCI file's header omits the leading row-id column: data rows are one field longer
            shift = len(first) - len(header)
            idx = {name: i + shift for i, name in enumerate(header)}
            for row in [first, *rd]:
                xs.append([float(row[idx[k]]) for k in FEATURES])
                ys.append(int(row[idx["Occupancy"]]))
        return np.asarray(xs, np.float32), np.asarray(ys, np.int64), path
    if os.path.exists(IN_TREE):
        with np.load(IN_TREE) as z:
            assert list(z["features"]) == FEATURES
            return z["x"].astype(np.float32), z["y"].astype(np.int64), "uci-occupancy (in-tree npz)"
    x, y = occupancy_like()
    return x, y, "synthetic"


def split_data(path: Optional[str] = None, clients_num: int = 20, *, test_size: float = 0.25,
               seed: int = 42):
    """-> (train shards, (X_test, y_test), source).  Mirrors ``split_data`` (M:33-49)."""
    x, y, src = load_table(path)
    rng = np.random.default_rng(seed)
    perm = rng.permutation(len(x))
    n_test = int(round(len(x) * test_size))
    te, tr = perm[:n_test], perm[n_test:]
    shards = []
    for xi, yi in zip(np.array_split(x[tr], clients_num), np.array_split(y[tr], clients_num)):
        shards.append(Shard(torch.from_numpy(xi.copy()), torch.from_numpy(yi.copy()), 2))
    test = Shard(torch.from_numpy(x[te].copy()), torch.from_numpy(y[te].copy()), 2)
    return shards, test, src


def one_hot(y: torch.Tensor, n: int = 2) -> torch.Tensor:
    """[1-y, y] encoding used by the reference (M:41-42)."""
    return torch.nn.functional.one_hot(y.long(), n).float()
