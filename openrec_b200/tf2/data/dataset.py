"""``Dataset``: the sampler entry points of the reference's data API (openrec/tf2/data/dataset.py:7-176), organised as
one table of *streams*.  A stream is an endless (or, for evaluation, finite) generator of feed dicts drawn from a
``_DataStore``; ``Dataset.<mode>(...)`` wraps the stream of that name in the spawn batcher.  The order of the calls
into ``random`` / the datastore inside every stream is what makes a seeded run reproduce the reference's sequence
(tests/golden/sampler.npz holds sequences recorded from the reference's own code)."""
from __future__ import annotations

import random

import numpy as np

from . import _DataStore, _ParallelDataset


class _Streams:
    """Feed-dict generators.  Every method takes the datastore first; extra arguments follow the feed spec below."""

    @staticmethod
    def pairwise(store):
        # dataset.py:7-16 -- next record of the epoch permutation, then ONE rejection-sampled negative for its user
        while True:
            record = store.next_random_record()
            u = record["user_id"]
            negative = store.sample_negative_items(u)[0]
            yield dict(user_id=u, p_item_id=record["item_id"], n_item_id=negative)

    @staticmethod
    def stratified(store, pos_ratio):
        # dataset.py:18-34 -- a coin per sample: observed record (label 1) or a uniformly drawn unobserved pair (label 0)
        users, items = store.total_users(), store.total_items()
        while True:
            take_positive = random.random() <= pos_ratio
            if take_positive:
                record = store.next_random_record()
                u, i = record["user_id"], record["item_id"]
            else:
                u, i = random.randint(0, users - 1), random.randint(0, items - 1)
                while store.is_positive(u, i):
                    u, i = random.randint(0, users - 1), random.randint(0, items - 1)
            yield dict(user_id=u, item_id=i, label=1.0 if take_positive else 0.0)

    @staticmethod
    def per_positive(store, pos_ratio):
        # dataset.py:36-58 -- every observed record, then int((1-r)/r) items drawn without replacement (minus the positive)
        quota = int((1 - pos_ratio) / pos_ratio)
        catalogue = range(store.total_items())
        while True:
            record = store.next_random_record()
            u, positive = record["user_id"], record["item_id"]
            yield dict(user_id=u, item_id=positive, label=1.0)
            candidates = [i for i in random.sample(catalogue, k=quota + 1) if i != positive]
            for i in candidates[:quota]:
                yield dict(user_id=u, item_id=i, label=0.0)

    @staticmethod
    def evaluation(store, exclude):
        # dataset.py:60-85 -- one row per warm user: which items count as hits, which are left out of the ranking
        size = store.total_items()
        only_listed_negatives = store.contain_negatives()
        for u in store.warm_users():
            liked = store.get_positive_items(u)
            hits = np.zeros(size, dtype=np.bool_)
            hits[liked] = True
            hidden = np.full(size, only_listed_negatives, dtype=np.bool_)
            if only_listed_negatives:
                hidden[liked] = False
                hidden[store.get_negative_items(u)] = False
            for other in exclude:
                hidden[other.datastore.get_positive_items(u)] = True
            yield dict(user_id=u, pos_mask=hits, excl_mask=hidden)


def _feed(scalars, vectors=(), width=0):
    """(dtypes, shapes) of a feed dict: ``scalars`` maps key -> dtype, ``vectors`` are bool masks of ``width``."""
    types = dict(scalars, **{k: "bool" for k in vectors})
    shapes = {k: [] for k in scalars}
    shapes.update({k: [width] for k in vectors})
    return types, shapes


_TRIPLET = {"user_id": "int32", "p_item_id": "int32", "n_item_id": "int32"}
_LABELLED = {"user_id": "int32", "item_id": "int32", "label": "float32"}


class Dataset:
    def __init__(self, raw_data, total_users, total_items, implicit_negative=True, num_negatives=None, seed=None,
                 sortby=None, asc=True, name=None):
        self.datastore = _DataStore(raw_data=raw_data, total_users=total_users, total_items=total_items,
                                    implicit_negative=implicit_negative, num_negatives=num_negatives, seed=seed,
                                    sortby=sortby, name=name, asc=asc)

    def _batched(self, stream, extra, feed, batch_size, workers, take=None):
        types, shapes = feed
        return _ParallelDataset(generator=stream, generator_params=(self.datastore,) + tuple(extra),
                                output_types=types, output_shapes=shapes, batch_size=batch_size,
                                num_parallel_calls=workers, take=take)

    def pairwise(self, batch_size, num_parallel_calls=1, take=None):
        return self._batched(_Streams.pairwise, (), _feed(_TRIPLET), batch_size, num_parallel_calls, take)

    def stratified_pointwise(self, batch_size, pos_ratio=0.5, num_parallel_calls=1, take=None):
        return self._batched(_Streams.stratified, (pos_ratio,), _feed(_LABELLED), batch_size, num_parallel_calls, take)

    def per_pos_stratified_pointwise(self, batch_size, pos_ratio=0.5, num_parallel_calls=1, take=None):
        return self._batched(_Streams.per_positive, (pos_ratio,), _feed(_LABELLED), batch_size, num_parallel_calls,
                             take)

    def evaluation(self, batch_size, excl_datasets=[]):
        feed = _feed({"user_id": "int32"}, ("pos_mask", "excl_mask"), self.datastore.total_items())
        return self._batched(_Streams.evaluation, (excl_datasets,), feed, batch_size, 1)
