"""Interaction index + multi-process batcher behind ``Dataset`` (reference:
openrec/tf2/data/utils.py:6-214).  numpy/stdlib only, so spawned workers start without torch.

Behavioural contract kept from the reference: structured-ndarray input else TypeError (utils.py:13-16);
``random.seed(seed)`` in the parent (utils.py:12); epoch-permutation record stream (utils.py:82-87);
uniform rejection sampling of negatives (utils.py:102-116) -- the in-process generators consume the
``random`` stream in the reference's order, so seeded sequences are identical (tests/golden/sampler.npz);
spawn-context daemon workers feeding a bounded queue, ``None`` sentinel, ``take`` (utils.py:164-214).
Deviation: ``random.sample`` on a set raises TypeError on Python >= 3.11 in the reference;
here the set is materialised first.
"""
from __future__ import annotations

import multiprocessing as mp
import random
import sys

import numpy as np


class _DataStore(object):
    def __init__(self, raw_data, total_users, total_items, implicit_negative=True, num_negatives=None, seed=None,
                 sortby=None, asc=True, name=None):
        self.name = name
        random.seed(seed)
        if type(raw_data) != np.ndarray:
            raise TypeError("Unsupported data input schema. Please use structured numpy array.")
        self._raw_data = raw_data
        self._rand_ids = []
        self._total_users, self._total_items = total_users, total_items
        self._sortby = sortby
        self._implicit_negative, self._num_negatives = implicit_negative, num_negatives
        store = self._index_store = {"positive": {}}

        def put(kind, user, item, ind):
            store[kind].setdefault(user, {})[item] = ind   # last record of a duplicated pair wins

        if implicit_negative:
            for ind, rec in enumerate(raw_data):
                put("positive", rec["user_id"], rec["item_id"], ind)
            if num_negatives is not None:
                store["negative"] = {}
                for user, pos in store["positive"].items():
                    chosen = store["negative"][user] = {}
                    for item in np.random.permutation(total_items):
                        if item not in pos:
                            chosen[item] = None
                        if len(chosen) == num_negatives:
                            break
        else:
            store["negative"] = {}
            for ind, rec in enumerate(raw_data):
                put("positive" if rec["label"] > 0 else "negative", rec["user_id"], rec["item_id"], ind)
        store["positive_sets"] = {u: set(d) for u, d in store["positive"].items()}
        if "negative" in store:
            store["negative_sets"] = {u: set(d) for u, d in store["negative"].items()}
        if sortby is not None:
            store["positive_sorts"] = {
                u: sorted(items, key=lambda it, u=u: raw_data[store["positive"][u][it]][sortby], reverse=not asc)
                for u, items in ((u, list(s)) for u, s in store["positive_sets"].items())}

    def contain_negatives(self):
        return not (self._implicit_negative and self._num_negatives is None)

    def next_random_record(self):
        """Next record of a per-epoch random permutation."""
        if not self._rand_ids:
            self._rand_ids = list(range(len(self._raw_data)))
            random.shuffle(self._rand_ids)
        return self._raw_data[self._rand_ids.pop()]

    def is_positive(self, user_id, item_id):
        return item_id in self._index_store["positive"].get(user_id, ())

    def sample_positive_items(self, user_id, num_samples=1):
        s = self._index_store["positive_sets"].get(user_id)
        return random.sample(list(s), num_samples) if s is not None else []

    def sample_negative_items(self, user_id, num_samples=1):
        neg = self._index_store.get("negative_sets")
        if neg is not None:
            s = neg.get(user_id)
            return random.sample(list(s), num_samples) if s is not None else []
        pos = self._index_store["positive_sets"].get(user_id, ())
        picked = set()
        cand = random.randint(0, self._total_items - 1)
        while len(picked) < num_samples:   # one extra draw after every test, as the reference does
            if cand not in pos:
                picked.add(cand)
            cand = random.randint(0, self._total_items - 1)
        return list(picked)

    def get_positive_items(self, user_id, sort=False):
        s = self._index_store["positive_sets"].get(user_id)
        if s is None:
            return []
        if sort:
            assert self._sortby is not None, "sortby key is not specified."
            return self._index_store["positive_sorts"][user_id]
        return list(s)

    def get_negative_items(self, user_id):
        neg = self._index_store.get("negative_sets")
        if neg is not None:
            return list(neg.get(user_id, ()))
        pos = self._index_store["positive_sets"][user_id]
        return [i for i in range(self._total_items) if i not in pos]

    def warm_users(self, threshold=1):
        return [u for u, d in self._index_store["positive"].items() if len(d) >= threshold]

    def total_users(self):
        return self._total_users

    def total_items(self):
        return self._total_items

    def total_records(self):
        return len(self._raw_data)


def _process(q, generator, generator_params, np_dtypes, batch_size):
    """Worker: run the generator, emit dict-of-ndarray batches, then the None sentinel."""
    keys = list(np_dtypes)
    cols = {k: [] for k in keys}
    n = 0

    def flush():
        q.put({k: np.asarray(cols[k], dtype=np_dtypes[k]) for k in keys})

    for sample in generator(*generator_params):
        for k in sample:
            cols[k].append(sample[k])
        n += 1
        if n == batch_size:
            flush()
            cols = {k: [] for k in keys}
            n = 0
    if n > 0:
        flush()
    q.put(None)


class _hidden_main:
    """The example scripts have no ``if __name__ == '__main__'`` guard; with the spawn start method a
    child would re-execute the script's top level.  Hide the main module's file/spec while starting
    workers so children import only this module."""

    def __enter__(self):
        self.main = sys.modules.get("__main__")
        self.saved = {}
        for attr in ("__file__", "__spec__"):
            if self.main is not None and getattr(self.main, attr, None) is not None:
                self.saved[attr] = getattr(self.main, attr)
                if attr == "__file__":
                    delattr(self.main, attr)
                else:
                    setattr(self.main, attr, None)

    def __exit__(self, *exc):
        for attr, v in self.saved.items():
            setattr(self.main, attr, v)
        return False


_NP_OF = {"int32": np.int32, "int64": np.int64, "float32": np.float32, "bool": np.bool_}


def _np_dtype(t):
    name = str(t).split(".")[-1]   # torch.int32 / tf.int32 -> 'int32'
    return _NP_OF[name]


class _ParallelDataset:
    def __init__(self, generator, generator_params, output_types, output_shapes, num_parallel_calls, batch_size,
                 take):
        ctx = mp.get_context("spawn")
        self._q = ctx.Queue(maxsize=num_parallel_calls)
        self._output_types = output_types
        self._take = take
        self._count = 0
        np_dtypes = {k: _np_dtype(output_types[k]) for k in output_shapes}
        self._p_list = []
        with _hidden_main():
            for _ in range(num_parallel_calls):
                p = ctx.Process(target=_process, args=(self._q, generator, generator_params, np_dtypes, batch_size))
                p.daemon = True
                p.start()
                self._p_list.append(p)

    def __iter__(self):
        return self

    def __next__(self):
        if self._take is not None and self._count >= self._take:
            raise StopIteration()
        batch = self._q.get()
        if batch is None:
            raise StopIteration()
        self._count += 1
        import torch   # parent only: stage through pinned memory to the device
        from ...tfshim.core import convert
        return {k: convert(v, getattr(torch, str(self._output_types[k]).split(".")[-1]))
                for k, v in batch.items()}
