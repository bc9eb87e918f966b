"""GPU: the reference-facing Python surface (openrec.tf2 + the tensorflow step protocol) against the
oracle -- same weights and ids injected on both sides."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import openrec_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tf():
    sys.path.insert(0, os.path.join(ROOT, "compat"))
    import tensorflow
    return tensorflow


def snapshot(model, names=("user_latent_factor", "item_latent_factor", "item_bias")):
    return [getattr(model, n).embeddings.numpy().astype(np.float64) for n in names]


def close(a, b, atol=1e-5, rtol=1e-5):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), atol=atol, rtol=rtol)


def make_ids(rng, U, I, B):
    return (rng.integers(0, U, B).astype(np.int32), rng.integers(0, I, B).astype(np.int32),
            rng.integers(0, I, B).astype(np.int32))


def train_step(tf, model, optimizer, *batch, objective=None):
    with tf.GradientTape() as tape:
        out = model(*batch)
    target = out if objective is None else objective(*out)
    grads = tape.gradient(target, model.trainable_variables)
    optimizer.apply_gradients(zip(grads, model.trainable_variables))
    return out


@pytest.mark.parametrize("optname", ["adam", "adagrad", "sgd", "lazy_adam"])
def test_bpr_protocol_matches_oracle(tf, optname):
    from openrec.tf2.recommenders import BPR
    rng = np.random.default_rng(11)
    U, I, D, B = 700, 900, 50, 1000   # the example's D=50 (generic-dim kernel)
    model = BPR(dim_user_embed=D, dim_item_embed=D, total_users=U, total_items=I)
    assert [tuple(v.shape) for v in model.trainable_variables] == [(U, D), (I, D), (I, 1)]
    opt = {"adam": tf.keras.optimizers.Adam, "adagrad": tf.keras.optimizers.Adagrad,
           "sgd": tf.keras.optimizers.SGD, "lazy_adam": tf.keras.optimizers.LazyAdam}[optname]()
    kind = {"adam": O.OPT_ADAM_DENSE, "adagrad": O.OPT_ADAGRAD, "sgd": O.OPT_SGD, "lazy_adam": O.OPT_ADAM_LAZY}[optname]
    user, item, bias = snapshot(model)
    assert np.abs(user).max() <= 0.05 and abs(user.std() - 0.1 / 12 ** 0.5) < 1e-3   # keras 'uniform'
    if kind == O.OPT_ADAGRAD:
        st = {k: (np.full_like(v, 0.1), None) for k, v in zip(("user", "item", "bias"), (user, item, bias))}
    else:
        st = {k: (np.zeros_like(v), np.zeros_like(v)) for k, v in zip(("user", "item", "bias"), (user, item, bias))}
    lr = opt.learning_rate
    for step in (1, 2, 3):
        uid, pid, nid = make_ids(rng, U, I, B)
        loss, l2 = train_step(tf, model, opt, tf.constant(uid), tf.constant(pid), tf.constant(nid))
        rl, rl2 = O.pairwise_train_step("bpr", user, item, bias, uid, pid, nid, kind, st, step, lr)
        close(loss.numpy(), rl, rtol=2e-5), close(l2.numpy(), rl2, rtol=2e-5)
        for got, want in zip(snapshot(model), (user, item, bias)):
            close(got, want)
    assert opt.iterations == 3


def test_lazy_scalars_and_weighted_objective(tf):
    from openrec.tf2.recommenders import BPR
    rng = np.random.default_rng(12)
    U, I, D, B = 200, 300, 64, 256
    model = BPR(D, D, U, I)
    user, item, bias = snapshot(model)
    uid, pid, nid = make_ids(rng, U, I, B)
    # no tape: reading the lazy scalars runs the forward-only kernel
    loss, l2 = model(uid, pid, nid)
    rl, rl2 = O.bpr_forward(user, item, bias, uid, pid, nid)
    close(float(loss), rl), close(float(l2), rl2, rtol=2e-5)
    for got, want in zip(snapshot(model), (user, item, bias)):
        assert np.array_equal(got, want)       # forward must not touch the tables
    # objective loss + 0.1*l2, read BEFORE apply_gradients
    opt = tf.keras.optimizers.SGD(learning_rate=0.5)
    with tf.GradientTape() as tape:
        loss, l2 = model(user_id=uid, p_item_id=pid, n_item_id=nid)     # keyword call (bpr_citeulike.py:53)
        obj = loss + 0.1 * l2
    close(float(obj), rl + 0.1 * rl2, rtol=2e-5)
    grads = tape.gradient(obj, model.trainable_variables)
    gr = O.bpr_grads(user, item, bias, uid, pid, nid, 1.0, 0.1)
    # materialised IndexedSlices (un-deduplicated, p||n concatenated)
    assert np.array_equal(grads[1].indices.numpy(), np.concatenate([pid, nid]))
    close(grads[0].values.numpy(), gr["user"][1]), close(grads[1].values.numpy(), gr["item"][1])
    close(grads[2].values.numpy(), gr["bias"][1].reshape(-1, 1))
    opt.apply_gradients(zip(grads, model.trainable_variables))
    for name, var in (("user", user), ("item", item), ("bias", bias)):
        idx, val = gr[name]
        O.sgd_sparse(var, idx, val.reshape(len(idx), -1), 0.5)
    for got, want in zip(snapshot(model), (user, item, bias)):
        close(got, want)
    with pytest.raises(RuntimeError):
        opt.apply_gradients(zip(grads, model.trainable_variables))      # a node steps once
    with pytest.raises(NotImplementedError):
        with tf.GradientTape() as tape:
            out = model(uid, pid, nid)
        g = tape.gradient(out, model.trainable_variables)
        opt.apply_gradients(zip(g[:2], model.trainable_variables[:2]))  # partial sets are not fused


def test_ucml_step_censor_inference(tf):
    from openrec.tf2.recommenders import UCML
    rng = np.random.default_rng(13)
    U, I, D, B = 300, 400, 32, 500
    model = UCML(D, D, U, I, margin=0.5)
    model.user_latent_factor.embeddings.assign(rng.uniform(-0.4, 0.4, (U, D)).astype(np.float32))
    model.item_latent_factor.embeddings.assign(rng.uniform(-0.4, 0.4, (I, D)).astype(np.float32))
    user, item, bias = snapshot(model)
    opt = tf.keras.optimizers.Adagrad(learning_rate=0.05)
    st = {k: (np.full_like(v, 0.1), None) for k, v in zip(("user", "item", "bias"), (user, item, bias))}
    uid, pid, nid = make_ids(rng, U, I, B)
    loss, l2 = train_step(tf, model, opt, uid, pid, nid)
    rl, rl2 = O.pairwise_train_step("ucml", user, item, bias, uid, pid, nid, O.OPT_ADAGRAD, st, 1, 0.05, margin=0.5)
    close(float(loss), rl, rtol=2e-5)
    model.censor_vec(uid, pid, nid)
    O.ucml_censor_vec(user, item, uid, pid, nid)
    for got, want in zip(snapshot(model), (user, item, bias)):
        close(got, want, atol=2e-5)
    close(model.inference(uid[:7]).numpy(), O.ucml_inference(user, item, bias, uid[:7]), atol=1e-4)


@pytest.mark.parametrize("kind", ["gmf", "wrmf"])
def test_pointwise_models(tf, kind):
    from openrec.tf2.recommenders import GMF, WRMF
    rng = np.random.default_rng(14)
    U, I, D, B = 250, 350, 64, 512
    model = GMF(D, D, U, I) if kind == "gmf" else WRMF(D, D, U, I, a=2.0, b=0.5)
    user, item, bias = snapshot(model)
    tv = model.trainable_variables
    w = tv[3].numpy().astype(np.float64) if kind == "gmf" else None
    assert len(tv) == (4 if kind == "gmf" else 3)
    if kind == "gmf":
        assert w.shape == (D, 1) and np.abs(w).max() <= (6 / (D + 1)) ** 0.5    # glorot uniform
    opt = tf.keras.optimizers.Adam()
    names = ("user", "item", "bias") + (("w",) if kind == "gmf" else ())
    st = {k: (np.zeros_like(v), np.zeros_like(v)) for k, v in zip(names, (user, item, bias, w))}
    for step in (1, 2):
        uid, iid, _ = make_ids(rng, U, I, B)
        label = (rng.random(B) < 0.3).astype(np.float32)
        loss, l2 = train_step(tf, model, opt, uid, iid, label)
        rl, rl2 = O.pointwise_train_step(kind, user, item, bias, w, uid, iid, label, O.OPT_ADAM_DENSE, st, step,
                                         0.001, 2.0, 0.5)
        close(float(loss), rl, rtol=2e-5), close(float(l2), rl2, rtol=2e-5)
        for got, want in zip(snapshot(model), (user, item, bias)):
            close(got, want)
        if kind == "gmf":
            close(tv[3].numpy(), w)
    ref = O.gmf_inference(user, item, bias, w, uid[:4]) if kind == "gmf" else O.dot_inference(user, item, bias, uid[:4])
    close(model.inference(uid[:4]).numpy(), ref)


def test_standalone_modules(tf):
    from openrec.tf2.modules import LatentFactor, PairwiseLogLoss, PointwiseMSELoss
    rng = np.random.default_rng(15)
    lf = LatentFactor(num_instances=50, dim=8)
    ids = rng.integers(0, 50, (3, 4))
    tab = lf.variables[0].numpy()
    assert np.array_equal(lf(ids).numpy(), tab[ids])                    # any id shape, bit exact
    z = LatentFactor(10, 4, zero_init=True)
    assert not z.variables[0].numpy().any()
    u, p, n = (rng.standard_normal((20, 8)).astype(np.float32) for _ in range(3))
    bp, bn = rng.standard_normal((20, 1)).astype(np.float32), rng.standard_normal((20, 1)).astype(np.float32)
    close(float(PairwiseLogLoss()(u, p, n, bp, bn)), O.pairwise_log_loss(u.astype(np.float64), p, n, bp, bn)[0])
    close(float(PairwiseLogLoss()(user_vec=u, p_item_vec=p, n_item_vec=n)),
          O.pairwise_log_loss(u.astype(np.float64), p, n)[0])
    lab = (rng.random(20) < 0.5).astype(np.float32)
    close(float(PointwiseMSELoss(a=2.0, b=0.5, sigmoid=True)(u, p, bp, lab)),
          O.pointwise_mse_loss(u.astype(np.float64), p, bp, lab.astype(np.float64), 2.0, 0.5, True)[0], rtol=2e-5)


def test_keras_metrics_and_tf_data(tf):
    m = tf.keras.metrics.Mean()
    m.update_state((tf.constant(2.0), tf.constant(4.0)))     # tuple averaged element-wise
    m.update_state(6.0)
    assert abs(float(m.result().numpy()) - 4.0) < 1e-6
    m.reset_states()
    rng = np.random.default_rng(16)
    y = (rng.random(4000) < 0.3).astype(np.float32)
    p = np.clip(y * 0.3 + rng.random(4000) * 0.7, 0, 1).astype(np.float32)
    auc = tf.keras.metrics.AUC()
    auc.update_state(y_true=y[:2000], y_pred=p[:2000])
    auc.update_state(y_true=y[2000:], y_pred=p[2000:])
    order = np.argsort(p)
    ranks = np.empty(len(p)); ranks[order] = np.arange(1, len(p) + 1)
    exact = (ranks[y > 0].sum() - (y > 0).sum() * ((y > 0).sum() + 1) / 2) / ((y > 0).sum() * (y == 0).sum())
    assert abs(float(auc.result().numpy()) - exact) < 5e-3       # 200-threshold Riemann sum
    data = {"a": np.arange(10, dtype=np.float32).reshape(10, 1), "b": np.arange(10, dtype=np.int32)}
    batches = list(tf.data.Dataset.from_tensor_slices(data).batch(4).prefetch(1).shuffle(20, seed=0))
    assert sorted(len(b["b"]) for b in batches) == [2, 4, 4]
    assert sorted(np.concatenate([b["b"].numpy() for b in batches]).tolist()) == list(range(10))
    assert all(np.array_equal(b["a"].numpy()[:, 0], b["b"].numpy()) for b in batches)


def test_example_flow_end_to_end():
    """Sampler workers -> fused steps (keras Adam, exact) -> evaluation -> metrics, as a user runs it."""
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "compat"), ROOT]))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "examples", "bpr_synthetic.py"), "--iters", "21",
                        "--users", "800", "--items", "1200", "--records", "20000", "--batch", "500"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("Iter:")]
    assert len(lines) == 3, r.stdout
    aucs = [float(l.split("AUC:")[1].split(",")[0]) for l in lines]
    assert all(0.0 <= a <= 1.0 for a in aucs)


def _sampler_dataset(seed=17, U=300, I=500, n=4000):
    from openrec.tf2.data import Dataset
    rng = np.random.default_rng(seed)
    pairs = np.unique(np.stack([rng.integers(0, U, n), rng.integers(0, I, n)], 1), axis=0)
    raw = np.empty(len(pairs), dtype=[("user_id", np.int32), ("item_id", np.int32)])
    raw["user_id"], raw["item_id"] = pairs[:, 0], pairs[:, 1]
    return Dataset(raw_data=raw, total_users=U, total_items=I), set(map(tuple, pairs.tolist())), len(raw), U, I


def test_device_sampler_semantics(tf):
    """DevicePairwiseSampler: every record exactly once per epoch -- batches that straddle the end of an epoch included,
    nothing dropped (data/utils.py:82-87) -- and negatives are never positive for the user."""
    from openrec.tf2.data import DevicePairwiseSampler
    ds, pos, n, U, I = _sampler_dataset()
    B = 500                                   # n is not a multiple of B: batches cross the epoch boundaries
    assert n % B != 0
    n_batches = (3 * n) // B
    it = DevicePairwiseSampler(ds, batch_size=B, take=n_batches, seed=3)
    stream = []
    for k, b in enumerate(it):
        u, p, q = (b[x].numpy() for x in ("user_id", "p_item_id", "n_item_id"))
        assert u.dtype == np.int32 and u.shape == (B,)
        assert all((int(a), int(c)) in pos for a, c in zip(u, p))          # positives are records
        assert not any((int(a), int(c)) in pos for a, c in zip(u, q))      # negatives are never positives
        assert q.min() >= 0 and q.max() < I
        stream += list(zip(u.tolist(), p.tolist()))
    assert k == n_batches - 1
    for e in range(len(stream) // n):                                      # every complete epoch = every record once
        assert set(stream[e * n:(e + 1) * n]) == pos
    assert len(set(stream[:n])) == n and stream[:n] != stream[n:2 * n]     # and a fresh permutation per epoch


def test_device_stratified_sampler(tf):
    """DeviceStratifiedSampler (dataset.py:18-34): label 1 = observed records in permutation order (each exactly once per
    epoch, across batch boundaries), label 0 = unobserved pairs, fraction of positives ~ pos_ratio."""
    from openrec.tf2.data import DeviceStratifiedSampler
    ds, pos, n, U, I = _sampler_dataset(seed=18)
    B, ratio = 700, 0.3
    it = DeviceStratifiedSampler(ds, batch_size=B, pos_ratio=ratio, take=40, seed=5)
    positives, n_lab1, total = [], 0, 0
    for b in it:
        u, i, lab = (b[x].numpy() for x in ("user_id", "item_id", "label"))
        assert lab.dtype == np.float32 and set(np.unique(lab).tolist()) <= {0.0, 1.0}
        for a, c, l in zip(u.tolist(), i.tolist(), lab.tolist()):
            assert ((a, c) in pos) == (l == 1.0)
        positives += [(a, c) for a, c, l in zip(u.tolist(), i.tolist(), lab.tolist()) if l == 1.0]
        n_lab1 += int(lab.sum())
        total += B
    assert abs(n_lab1 / total - ratio) < 0.03
    assert len(positives) > n                                              # more than one epoch of records was consumed
    assert set(positives[:n]) == pos and len(set(positives[:n])) == n      # the first epoch: every record exactly once


def test_device_per_positive_sampler(tf):
    """DevicePerPositiveSampler (dataset.py:36-58): the stream record, quota distinct other items, record, ... cut into
    batches at arbitrary positions; group content does not depend on where the cut falls."""
    from openrec.tf2.data import DevicePerPositiveSampler
    ds, pos, n, U, I = _sampler_dataset(seed=19, n=1500)
    ratio, quota = 0.2, 4

    def stream(B, batches):
        out = []
        for b in DevicePerPositiveSampler(ds, batch_size=B, pos_ratio=ratio, take=batches, seed=7):
            out += list(zip(b["user_id"].numpy().tolist(), b["item_id"].numpy().tolist(), b["label"].numpy().tolist()))
        return out
    s1 = stream(333, 30)                       # 333 is not a multiple of the group size 5
    s2 = stream(999, 10)
    assert s1 == s2                            # same stream however it is batched
    g = quota + 1
    recs = []
    for k in range(len(s1) // g):
        grp = s1[k * g:(k + 1) * g]
        u, p, l = grp[0]
        assert l == 1.0 and (u, p) in pos
        negs = [x[1] for x in grp[1:]]
        assert all(x[0] == u and x[2] == 0.0 for x in grp[1:])
        assert len(set(negs)) == quota and p not in negs and min(negs) >= 0 and max(negs) < I
        recs.append((u, p))
    assert len(recs) > n and set(recs[:n]) == pos and len(set(recs[:n])) == n


def test_checkpoint_roundtrip(tf, tmp_path):
    from openrec.tf2.recommenders import BPR
    from openrec_b200.tf2 import checkpoint
    rng = np.random.default_rng(18)
    U, I, D, B = 100, 150, 64, 256
    m1 = BPR(D, D, U, I)
    o1 = tf.keras.optimizers.Adagrad(learning_rate=0.05)
    ids = make_ids(rng, U, I, B)
    train_step(tf, m1, o1, *ids)
    checkpoint.save(tmp_path / "ck", m1, o1)
    m2 = BPR(D, D, U, I)
    o2 = tf.keras.optimizers.Adagrad(learning_rate=0.05)
    checkpoint.load(tmp_path / "ck", m2, o2)
    ids2 = make_ids(rng, U, I, B)
    l1 = train_step(tf, m1, o1, *ids2)
    l2 = train_step(tf, m2, o2, *ids2)
    assert abs(float(l1[0]) - float(l2[0])) <= 1e-6 * abs(float(l1[0]))
    for a, b in zip(snapshot(m1), snapshot(m2)):
        np.testing.assert_allclose(a, b, rtol=1e-5, atol=1e-7)   # fp32 RED order on shared rows is not fixed run to run


def test_sharded_class_surface_single_rank(tf, tmp_path):
    """openrec.tf2.recommenders.ShardedBPR keeps the BPR constructor and step protocol on row-sharded tables (here a
    world of one rank, so the whole flow -- route, request, serve, compute, apply, tail, flag words -- runs on this GPU):
    same losses and tables as the single-GPU BPR started from the same weights; shard checkpoint round trip."""
    import torch.distributed as dist
    from openrec.tf2.recommenders import BPR, ShardedBPR
    from openrec_b200.tf2 import checkpoint
    created = False
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{29700 + os.getpid() % 200}", rank=0, world_size=1)
        created = True
    try:
        rng = np.random.default_rng(21)
        U, I, D, B = 211, 307, 64, 512
        ref = BPR(D, D, U, I)
        sh = ShardedBPR(D, D, U, I)
        for a, b in zip(sh.variables, ref.variables):
            a.assign(b.numpy())
        o_ref = tf.keras.optimizers.Adagrad(learning_rate=0.05)
        o_sh = tf.keras.optimizers.Adagrad(learning_rate=0.05)
        for step in range(3):
            ids = make_ids(rng, U, I, B)
            l_ref = train_step(tf, ref, o_ref, *ids)
            l_sh = train_step(tf, sh, o_sh, *ids)
            np.testing.assert_allclose(float(l_sh[0]), float(l_ref[0]), rtol=2e-5)
            np.testing.assert_allclose(float(l_sh[1]), float(l_ref[1]), rtol=2e-5)
        sh.check()
        for a, b in zip(sh.variables, ref.variables):
            np.testing.assert_allclose(a.numpy(), b.numpy(), atol=1e-6)
        checkpoint.save(tmp_path / "shard0", sh, o_sh)
        sh2 = ShardedBPR(D, D, U, I, seed=5)
        o2 = tf.keras.optimizers.Adagrad(learning_rate=0.05)
        checkpoint.load(tmp_path / "shard0", sh2, o2)
        ids = make_ids(rng, U, I, B)
        l1, l2 = train_step(tf, sh, o_sh, *ids), train_step(tf, sh2, o2, *ids)
        np.testing.assert_allclose(float(l1[0]), float(l2[0]), rtol=1e-6)
        with pytest.raises(NotImplementedError):
            train_step(tf, ShardedBPR(D, D, U, I), tf.keras.optimizers.Adam(), *ids)
    finally:
        if created:
            dist.destroy_process_group()
