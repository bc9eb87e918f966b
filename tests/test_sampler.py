"""CPU: the sampler API (openrec_b200.tf2.data) against sequences recorded from the reference's own
generator code (tests/golden/sampler.npz) + host-logic properties."""
import os

import numpy as np
import pytest

from openrec_b200.tf2.data import Dataset, _DataStore
from openrec_b200.tf2.data import dataset as D


@pytest.fixture(scope="module")
def g(golden_dir):
    return dict(np.load(os.path.join(golden_dir, "sampler.npz")))


def _raw(g):
    raw = np.empty(len(g["raw_user"]), dtype=[("user_id", np.int32), ("item_id", np.int32)])
    raw["user_id"], raw["item_id"] = g["raw_user"], g["raw_item"]
    return raw


def _take(gen, n):
    rows = []
    for _ in range(n):
        d = next(gen)
        rows.append([float(d[k]) for k in sorted(d)])
    return np.array(rows)


def test_pairwise_stream_bit_exact(g):
    ds = _DataStore(raw_data=_raw(g), total_users=int(g["U"]), total_items=int(g["I"]), seed=123)
    got = _take(D._Streams.pairwise(ds), len(g["pairwise"]))
    assert np.array_equal(got, g["pairwise"])          # same records, same negatives, same order
    pos = {}
    for u, i in zip(g["raw_user"], g["raw_item"]):
        pos.setdefault(int(u), set()).add(int(i))
    n, p, u = got[:, 0], got[:, 1], got[:, 2]
    assert all(int(ni) not in pos[int(ui)] for ni, ui in zip(n, u))      # negatives never positive
    assert all(int(pi) in pos[int(ui)] for pi, ui in zip(p, u))
    first_epoch = set(zip(u[:300].astype(int), p[:300].astype(int)))
    assert len(first_epoch) == 300                                       # each record exactly once per epoch


def test_pointwise_streams_bit_exact(g):
    ds = _DataStore(raw_data=_raw(g), total_users=int(g["U"]), total_items=int(g["I"]), seed=5)
    assert np.array_equal(_take(D._Streams.stratified(ds, 0.3), len(g["stratified"])), g["stratified"])
    ds = _DataStore(raw_data=_raw(g), total_users=int(g["U"]), total_items=int(g["I"]), seed=9)
    assert np.array_equal(_take(D._Streams.per_positive(ds, 0.2), len(g["per_pos"])), g["per_pos"])


def test_evaluation_generator(g):
    raw = _raw(g)
    tr = Dataset(raw_data=raw[:200], total_users=int(g["U"]), total_items=int(g["I"]), seed=1)
    va = _DataStore(raw_data=raw[200:], total_users=int(g["U"]), total_items=int(g["I"]), seed=1)
    ev = list(D._Streams.evaluation(va, [tr]))
    assert np.array_equal(np.array([e["user_id"] for e in ev], dtype=np.int32), g["eval_user"])
    assert np.array_equal(np.stack([e["pos_mask"] for e in ev]), g["eval_pos"])
    assert np.array_equal(np.stack([e["excl_mask"] for e in ev]), g["eval_excl"])


def test_datastore_contract(g):
    with pytest.raises(TypeError):
        _DataStore(raw_data=[(1, 2)], total_users=3, total_items=3)
    raw = _raw(g)
    ds = _DataStore(raw_data=raw, total_users=int(g["U"]), total_items=int(g["I"]), seed=0)
    assert ds.total_records() == len(raw) and not ds.contain_negatives()
    u, i = int(raw["user_id"][0]), int(raw["item_id"][0])
    assert ds.is_positive(u, i) and i in ds.get_positive_items(u)
    assert set(ds.get_negative_items(u)).isdisjoint(ds.get_positive_items(u))
    assert len(ds.sample_positive_items(u, 1)) == 1        # works on py>=3.11
    ds2 = _DataStore(raw_data=raw, total_users=int(g["U"]), total_items=int(g["I"]), num_negatives=5, seed=0)
    assert ds2.contain_negatives()
    negs = ds2.sample_negative_items(u, 3)
    assert len(negs) == 3 and set(negs) <= set(ds2.get_negative_items(u)) and len(ds2.get_negative_items(u)) == 5
    lab = np.empty(4, dtype=[("user_id", np.int32), ("item_id", np.int32), ("label", np.float32)])
    lab["user_id"], lab["item_id"], lab["label"] = [0, 0, 1, 1], [1, 2, 1, 3], [1, 0, 1, 0]
    ds3 = _DataStore(raw_data=lab, total_users=2, total_items=4, implicit_negative=False)
    assert ds3.get_positive_items(0) == [1] and ds3.get_negative_items(0) == [2] and ds3.contain_negatives()


def test_parallel_workers_batches_sentinel_take(g):
    """spawned workers -> bounded queue of dict-of-ndarray batches; remainder + None; take."""
    raw = _raw(g)
    U, I = int(g["U"]), int(g["I"])
    ds = Dataset(raw_data=raw, total_users=U, total_items=I, seed=3)
    it = ds.pairwise(batch_size=64, num_parallel_calls=2, take=3)
    pos = {}
    for u, i in zip(raw["user_id"], raw["item_id"]):
        pos.setdefault(int(u), set()).add(int(i))
    for _ in range(4):
        b = it._q.get(timeout=60)
        assert set(b) == {"user_id", "p_item_id", "n_item_id"}
        assert all(v.dtype == np.int32 and v.shape == (64,) for v in b.values())
        assert all(int(n) not in pos[int(u)] for n, u in zip(b["n_item_id"], b["user_id"]))
    it._count = 3
    with pytest.raises(StopIteration):
        next(it)
    ev = ds.evaluation(batch_size=16, excl_datasets=[])
    n_users, batches = len(ds.datastore.warm_users()), []
    while True:
        b = ev._q.get(timeout=60)
        if b is None:
            break
        batches.append(b)
    assert sum(len(b["user_id"]) for b in batches) == n_users
    assert [len(b["user_id"]) for b in batches[:-1]] == [16] * (len(batches) - 1)   # remainder batch last
    assert batches[0]["pos_mask"].dtype == np.bool_ and batches[0]["pos_mask"].shape[1] == I
    for p in it._p_list + ev._p_list:
        p.terminate()
