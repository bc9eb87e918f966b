"""``Dataset`` sampler API (reference: openrec/tf2/data/dataset.py:7-176)."""
from __future__ import annotations

import random

import numpy as np

from . import _DataStore, _ParallelDataset

_INT32, _FLOAT32, _BOOL = "int32", "float32", "bool"


def _pairwise_generator(datastore):
    """(user, positive item) from the epoch permutation + one rejection-sampled negative (dataset.py:7-16)."""
    while True:
        rec = datastore.next_random_record()
        user = rec["user_id"]
        yield {"user_id": user, "p_item_id": rec["item_id"],
               "n_item_id": datastore.sample_negative_items(user)[0]}


def _stratified_pointwise_generator(datastore, pos_ratio):
    """Bernoulli(pos_ratio) mix of observed records (label 1) and uniform non-positive pairs (label 0)
    (dataset.py:18-34)."""
    n_users, n_items = datastore.total_users(), datastore.total_items()
    while True:
        if random.random() <= pos_ratio:
            rec = datastore.next_random_record()
            yield {"user_id": rec["user_id"], "item_id": rec["item_id"], "label": 1.0}
            continue
        while True:
            user, item = random.randint(0, n_users - 1), random.randint(0, n_items - 1)
            if not datastore.is_positive(user, item):
                break
        yield {"user_id": user, "item_id": item, "label": 0.0}


def _per_pos_stratified_pointwise_generator(datastore, pos_ratio):
    """Each observed record followed by int((1-r)/r) sampled items != the positive (dataset.py:36-58)."""
    per_pos = int((1 - pos_ratio) / pos_ratio)
    while True:
        rec = datastore.next_random_record()
        user, pos_item = rec["user_id"], rec["item_id"]
        yield {"user_id": user, "item_id": pos_item, "label": 1.0}
        emitted = 0
        for item in random.sample(range(datastore.total_items()), k=per_pos + 1):
            if item == pos_item:
                continue
            yield {"user_id": user, "item_id": item, "label": 0.0}
            emitted += 1
            if emitted >= per_pos:
                break


def _evaluation_generator(datastore, excl_datasets):
    """Per warm user: positives mask and exclusion mask over the catalogue (dataset.py:60-85)."""
    n_items = datastore.total_items()
    for user in datastore.warm_users():
        positives = datastore.get_positive_items(user)
        pos_mask = np.zeros(n_items, dtype=np.bool_)
        pos_mask[positives] = True
        if datastore.contain_negatives():   # only listed negatives are evaluated
            excl_mask = np.ones(n_items, dtype=np.bool_)
            excl_mask[positives] = False
            excl_mask[datastore.get_negative_items(user)] = False
        else:
            excl_mask = np.zeros(n_items, dtype=np.bool_)
        seen = []
        for other in excl_datasets:
            seen += other.datastore.get_positive_items(user)
        excl_mask[seen] = True
        yield {"user_id": user, "pos_mask": pos_mask, "excl_mask": excl_mask}


class Dataset:
    def __init__(self, raw_data, total_users, total_items, implicit_negative=True, num_negatives=None, seed=None,
                 sortby=None, asc=True, name=None):
        self.datastore = _DataStore(raw_data=raw_data, total_users=total_users, total_items=total_items,
                                    implicit_negative=implicit_negative, num_negatives=num_negatives, seed=seed,
                                    sortby=sortby, name=name, asc=asc)

    def _build_dataset(self, generator, generator_params, output_types, output_shapes, batch_size,
                       num_parallel_calls, take=None):
        return _ParallelDataset(generator=generator, generator_params=generator_params, output_types=output_types,
                                output_shapes=output_shapes, batch_size=batch_size,
                                num_parallel_calls=num_parallel_calls, take=take)

    def pairwise(self, batch_size, num_parallel_calls=1, take=None):
        keys = ("user_id", "p_item_id", "n_item_id")
        return self._build_dataset(_pairwise_generator, (self.datastore,), {k: _INT32 for k in keys},
                                   {k: [] for k in keys}, batch_size, num_parallel_calls, take)

    def _pointwise(self, generator, batch_size, pos_ratio, num_parallel_calls, take):
        types = {"user_id": _INT32, "item_id": _INT32, "label": _FLOAT32}
        return self._build_dataset(generator, (self.datastore, pos_ratio), types, {k: [] for k in types},
                                   batch_size, num_parallel_calls, take)

    def stratified_pointwise(self, batch_size, pos_ratio=0.5, num_parallel_calls=1, take=None):
        return self._pointwise(_stratified_pointwise_generator, batch_size, pos_ratio, num_parallel_calls, take)

    def per_pos_stratified_pointwise(self, batch_size, pos_ratio=0.5, num_parallel_calls=1, take=None):
        return self._pointwise(_per_pos_stratified_pointwise_generator, batch_size, pos_ratio, num_parallel_calls,
                               take)

    def evaluation(self, batch_size, excl_datasets=[]):
        n = self.datastore.total_items()
        types = {"user_id": _INT32, "pos_mask": _BOOL, "excl_mask": _BOOL}
        shapes = {"user_id": [], "pos_mask": [n], "excl_mask": [n]}
        return self._build_dataset(_evaluation_generator, (self.datastore, excl_datasets), types, shapes,
                                   batch_size, 1)
