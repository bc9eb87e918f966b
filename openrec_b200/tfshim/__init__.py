"""Minimal ``tensorflow`` namespace for running tf2_examples/*.py unmodified on liborx.

Only the symbols the two examples and user-level glue need are provided; anything else raises
NotImplementedError (this is a step-protocol adaptor for the openrec.tf2 hot path, not a TensorFlow).
Put ``<repo>/compat`` on PYTHONPATH to expose it as ``import tensorflow``.
"""
from __future__ import annotations

import torch as _torch

from . import data, keras  # noqa: F401
from .core import (GradientTape, LazyScalar, SparseGrad, Tensor, Variable, bool_, convert, float32, float64,
                   function, int32, int64, uint8, unwrap)
from .keras.layers import set_seed as _set_seed

__version__ = "2.0.1-orx"
bool = bool_  # noqa: A001  (tf.bool)


def constant(value, dtype=None, shape=None, name=None):
    t = convert(value, dtype)
    return Tensor(t.t.reshape(shape)) if shape is not None else t


convert_to_tensor = constant


def cast(x, dtype):
    return Tensor(convert(x).t.to(dtype))


def reshape(x, shape):
    return Tensor(convert(x).t.reshape(tuple(int(s) for s in shape)))


def zeros(shape, dtype=float32):
    from .core import device
    return Tensor(_torch.zeros(tuple(int(s) for s in shape), dtype=dtype, device=device()))


def shape(x):
    return tuple(convert(x).t.shape)


def reduce_sum(x, axis=None, keepdims=False):
    t = convert(x).t
    return Tensor(t.sum() if axis is None else t.sum(dim=axis, keepdim=keepdims))


def reduce_mean(x, axis=None, keepdims=False):
    t = convert(x).t.to(_torch.float32)
    return Tensor(t.mean() if axis is None else t.mean(dim=axis, keepdim=keepdims))


class _Random:
    @staticmethod
    def set_seed(seed):
        _set_seed(seed)


random = _Random()


def __getattr__(name):
    raise NotImplementedError(
        f"tensorflow.{name} is not provided by the openrec_b200 shim (only the symbols used by "
        "openrec.tf2 and tf2_examples are; see openrec_b200/tfshim/__init__.py)")
