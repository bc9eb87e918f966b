"""Table / optimizer-slot checkpoint I/O.  The reference's tf2 path has none (``save_interval`` is
assigned and never used, tf2_examples/bpr_citeulike.py:16); 51 GB sharded tables need one to be practical.
Format: one ``.npz`` with ``var/<k>`` in ``model.variables`` order and their ``names``, plus ``slot<j>/<k>`` and
``iterations`` when an optimizer is given.  Row-sharded models: ``HomeRoutedPairwise.save_shard / load_shard``
(openrec_b200/sharded.py), one file per rank.

A Keras model creates its Dense layers during the first call, so an un-called DLRM / GMF does not have all of its
variables yet: saving or loading such a model would silently drop the MLP weights.  Both directions therefore insist
that the variable lists match (count, names, shapes); ``load(..., build=fn)`` lets the caller run one forward call
(``fn()``) first when the model is fresh."""
from __future__ import annotations

import numpy as np
import torch


def _unbuilt(model):
    """Names of sub-layers that have not created their variables yet (keras `built` flag)."""
    out = []
    for name, sub in vars(model).items():
        for layer in getattr(sub, "layers", [sub]):
            if hasattr(layer, "built") and not layer.built:
                out.append(f"{name}.{type(layer).__name__}")
    return out


def save(path, model, optimizer=None):
    pending = _unbuilt(model)
    if pending:
        raise ValueError(f"checkpoint.save: {pending} have no variables yet (call the model once first); a checkpoint "
                         "written now would silently lack them")
    variables = model.variables
    out = {f"var/{k}": v.numpy() for k, v in enumerate(variables)}
    out["names"] = np.array([v.name for v in variables])
    if optimizer is not None:
        out["iterations"] = np.int64(optimizer.iterations)
        for k, v in enumerate(variables):
            for j, s in enumerate(optimizer.slots_if_any(v)):
                if s is not None:
                    out[f"slot{j}/{k}"] = s.detach().cpu().numpy()
    np.savez(path, **out)


def load(path, model, optimizer=None, build=None):
    data = np.load(path if str(path).endswith(".npz") else str(path) + ".npz", allow_pickle=False)
    if build is not None and _unbuilt(model):
        build()
    n_saved = sum(1 for k in data.files if k.startswith("var/"))
    variables = model.variables
    pending = _unbuilt(model)
    if pending or n_saved != len(variables):
        raise ValueError(f"checkpoint holds {n_saved} variables, the model has {len(variables)}"
                         + (f" ({pending} not built yet: call the model once, or pass build=)" if pending else ""))
    names = [str(x) for x in data["names"]] if "names" in data.files else None
    for k, v in enumerate(variables):
        a = data[f"var/{k}"]
        if names is not None and names[k] != v.name:
            raise ValueError(f"checkpoint variable {k} is {names[k]!r}, the model's is {v.name!r}")
        if tuple(a.shape) != tuple(v.shape):
            raise ValueError(f"checkpoint variable {k} has shape {a.shape}, model expects {tuple(v.shape)}")
    for k, v in enumerate(variables):
        v.assign(data[f"var/{k}"])
    if optimizer is not None and "iterations" in data.files:
        optimizer.iterations = int(data["iterations"])
        for k, v in enumerate(variables):
            s = list(optimizer.slots(v))
            for j in range(2):
                if f"slot{j}/{k}" in data.files and s[j] is not None:
                    s[j].copy_(torch.from_numpy(data[f"slot{j}/{k}"]))
