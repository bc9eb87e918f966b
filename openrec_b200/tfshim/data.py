"""tensorflow.data.Dataset: from_tensor_slices(...).batch().prefetch().shuffle() as used by
tf2_examples/dlrm_criteo.py:18-29.  Host-side slicing; each yielded batch is staged through pinned
memory onto the device."""
from __future__ import annotations

import numpy as np

from .core import convert


class Dataset:
    def __init__(self, source, kind, arg=None):
        self._source, self._kind, self._arg = source, kind, arg

    @staticmethod
    def from_tensor_slices(tensors):
        if isinstance(tensors, dict):
            arrs = {k: np.asarray(v) for k, v in tensors.items()}
            n = {len(a) for a in arrs.values()}
            if len(n) != 1:
                raise ValueError("from_tensor_slices: all components must share dimension 0")
        else:
            arrs = np.asarray(tensors)
        return Dataset(arrs, "slices")

    def batch(self, batch_size, drop_remainder=False):
        return Dataset(self, "batch", (int(batch_size), bool(drop_remainder)))

    def prefetch(self, buffer_size):
        return Dataset(self, "prefetch", buffer_size)

    def shuffle(self, buffer_size, seed=None, reshuffle_each_iteration=True):
        return Dataset(self, "shuffle", (int(buffer_size), seed))

    def take(self, count):
        return Dataset(self, "take", int(count))

    def repeat(self, count=None):
        return Dataset(self, "repeat", count)

    # host-level iteration over numpy elements
    def _iter_host(self):
        k = self._kind
        if k == "slices":
            a = self._source
            if isinstance(a, dict):
                n = len(next(iter(a.values())))
                for i in range(n):
                    yield {key: v[i] for key, v in a.items()}
            else:
                yield from a
        elif k == "batch":
            bs, drop = self._arg
            src = self._source
            if src._kind == "slices":   # fast path: slice the arrays directly
                a = src._source
                n = len(next(iter(a.values()))) if isinstance(a, dict) else len(a)
                for s in range(0, n, bs):
                    if drop and s + bs > n:
                        break
                    yield {key: v[s:s + bs] for key, v in a.items()} if isinstance(a, dict) else a[s:s + bs]
            else:
                buf = []
                for e in src._iter_host():
                    buf.append(e)
                    if len(buf) == bs:
                        yield _stack(buf)
                        buf = []
                if buf and not drop:
                    yield _stack(buf)
        elif k == "prefetch":
            yield from self._source._iter_host()
        elif k == "shuffle":   # tf.data buffer shuffle: shuffles whatever the elements are (batches, Q11)
            size, seed = self._arg
            rng = np.random.default_rng(seed)
            buf = []
            for e in self._source._iter_host():
                if len(buf) < size:
                    buf.append(e)
                    continue
                j = int(rng.integers(0, size))
                yield buf[j]
                buf[j] = e
            rng.shuffle(buf)
            yield from buf
        elif k == "take":
            for i, e in enumerate(self._source._iter_host()):
                if i >= self._arg:
                    break
                yield e
        elif k == "repeat":
            c = 0
            while self._arg is None or c < self._arg:
                yield from self._source._iter_host()
                c += 1

    def __iter__(self):
        for e in self._iter_host():
            yield {k: convert(v) for k, v in e.items()} if isinstance(e, dict) else convert(e)


def _stack(elems):
    if isinstance(elems[0], dict):
        return {k: np.stack([e[k] for e in elems]) for k in elems[0]}
    return np.stack(elems)
