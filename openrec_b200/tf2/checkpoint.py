"""Table / optimizer-slot checkpoint I/O.  The reference's tf2 path has none (``save_interval`` is
assigned and never used, tf2_examples/bpr_citeulike.py:16); 51 GB sharded tables need one to be practical.
Format: one ``.npz`` with ``var/<k>`` in ``model.variables`` order, plus ``slot<j>/<k>`` and ``iterations`` when an
optimizer is given."""
from __future__ import annotations

import numpy as np
import torch


def save(path, model, optimizer=None):
    out = {f"var/{k}": v.numpy() for k, v in enumerate(model.variables)}
    out["names"] = np.array([v.name for v in model.variables])
    if optimizer is not None:
        out["iterations"] = np.int64(optimizer.iterations)
        for k, v in enumerate(model.variables):
            for j, s in enumerate(optimizer._slots.get(id(v), ())):
                if s is not None:
                    out[f"slot{j}/{k}"] = s.detach().cpu().numpy()
    np.savez(path, **out)


def load(path, model, optimizer=None):
    data = np.load(path if str(path).endswith(".npz") else str(path) + ".npz", allow_pickle=False)
    variables = model.variables
    for k, v in enumerate(variables):
        a = data[f"var/{k}"]
        if tuple(a.shape) != tuple(v.shape):
            raise ValueError(f"checkpoint variable {k} has shape {a.shape}, model expects {tuple(v.shape)}")
        v.assign(a)
    if optimizer is not None and "iterations" in data:
        optimizer.iterations = int(data["iterations"])
        for k, v in enumerate(variables):
            s = list(optimizer.slots(v))
            for j in range(2):
                if f"slot{j}/{k}" in data and s[j] is not None:
                    s[j].copy_(torch.from_numpy(data[f"slot{j}/{k}"]))
