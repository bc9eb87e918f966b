"""tf.keras.metrics.{Mean, AUC} as used by tf2_examples (bpr_citeulike.py:48,54,64-66;
dlrm_criteo.py:40,53,70).  Evaluation-side glue on device tensors, not the timed path."""
from __future__ import annotations

import torch

from ..core import LazyScalar, Tensor, convert, device, unwrap


def _flatten(values):
    if isinstance(values, (list, tuple)):
        out = []
        for v in values:
            out += _flatten(v)
        return out
    return [values]


class Mean:
    def __init__(self, name="mean", dtype=None):
        self.name = name
        self.reset_states()

    def reset_states(self):
        self._total = None
        self._count = 0

    reset_state = reset_states

    def update_state(self, values, sample_weight=None):
        if sample_weight is not None:
            raise NotImplementedError("Mean(sample_weight=...)")
        for v in _flatten(values):   # a tuple (loss, l2_loss) is averaged element-wise
            t = unwrap(v) if isinstance(v, (Tensor, LazyScalar)) or hasattr(v, "t") else convert(v).t
            t = t.to(torch.float32)
            s = t.sum() if t.dim() else t
            self._total = s if self._total is None else self._total + s
            self._count += max(t.numel(), 1)

    def result(self):
        if self._total is None:
            return Tensor(torch.zeros((), device=device()))
        return Tensor(self._total / float(self._count))


class AUC:
    """Keras AUC defaults: 200 thresholds, ROC curve, 'interpolation' (trapezoid) [TF-mem]."""

    def __init__(self, num_thresholds=200, curve="ROC", summation_method="interpolation", name="auc"):
        if curve != "ROC" or summation_method != "interpolation":
            raise NotImplementedError("only the Keras default ROC/interpolation AUC is provided")
        n = num_thresholds
        eps = 1e-7
        th = [0.0 - eps] + [(i + 1) / (n - 1) for i in range(n - 2)] + [1.0 + eps]
        self._th = torch.tensor(th, dtype=torch.float32, device=device())
        self.reset_states()

    def reset_states(self):
        z = torch.zeros_like(self._th, dtype=torch.float64)
        self._tp, self._fp, self._tn, self._fn = z.clone(), z.clone(), z.clone(), z.clone()

    reset_state = reset_states

    def update_state(self, y_true, y_pred, sample_weight=None):
        y = convert(y_true).t.reshape(-1).to(torch.float32) > 0.5
        p = convert(y_pred).t.reshape(-1).to(torch.float32)
        above = p.unsqueeze(0) > self._th.unsqueeze(1)          # [T, N]
        pos = y.unsqueeze(0)
        self._tp += (above & pos).sum(1)
        self._fp += (above & ~pos).sum(1)
        self._fn += (~above & pos).sum(1)
        self._tn += (~above & ~pos).sum(1)

    def result(self):
        def dnn(a, b):
            return torch.where(b > 0, a / b.clamp(min=1), torch.zeros_like(a))
        tpr = dnn(self._tp, self._tp + self._fn)
        fpr = dnn(self._fp, self._fp + self._tn)
        auc = ((fpr[:-1] - fpr[1:]) * (tpr[:-1] + tpr[1:]) / 2.0).sum()
        return Tensor(auc.to(torch.float32))
