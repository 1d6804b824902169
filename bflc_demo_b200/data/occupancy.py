"""The reference's data pipeline (python-sdk/main.py:32-49, component P2): read the UCI
Occupancy CSV, split 75/25 with a fixed seed, one-hot labels, contiguous IID shards via
``np.array_split``.  sklearn/pandas are not required: the split is a seeded permutation
(documented deviation: not bit-identical to ``train_test_split(random_state=42)``, same
proportions).  Falls back to a schema-compatible synthetic table when the CSV is absent."""
from __future__ import annotations

import csv
import os
from typing import List, Optional, Tuple

import numpy as np
import torch

from .synthetic import Shard, occupancy_like

FEATURES = ["Temperature", "Humidity", "Light", "CO2", "HumidityRatio"]  # M:35-36
DEFAULT_CSV = "/root/reference/python-sdk/data/datatraining.txt"


def load_table(path: Optional[str] = None) -> Tuple[np.ndarray, np.ndarray, str]:
    path = path or os.environ.get("BFLC_OCCUPANCY_CSV", DEFAULT_CSV)
    if path and os.path.exists(path):
        xs, ys = [], []
        with open(path, newline="") as f:
            rd = csv.reader(f)
            header = next(rd)
            # the file's header omits the leading row-id column
            cols = header if len(header) == 7 else header
            idx = {name.strip('"'): i + (1 if len(cols) == 7 else 0) for i, name in enumerate(cols)}
            for row in rd:
                xs.append([float(row[idx[k]]) for k in FEATURES])
                ys.append(int(row[idx["Occupancy"]]))
        return np.asarray(xs, np.float32), np.asarray(ys, np.int64), path
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
