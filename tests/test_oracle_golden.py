"""Pin oracle/openrec_oracle.py to the golden vectors recorded from the reference's own
Python code (tests/golden/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import openrec_oracle as O


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name)))


def _dense(rows, idx, val):
    out = np.zeros((rows,) + val.shape[1:], dtype=val.dtype)
    np.add.at(out, np.asarray(idx, dtype=np.int64), val)
    return out


@pytest.mark.parametrize("kind", ["bpr", "ucml"])
@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-12), (np.float32, 1e-5)])
def test_pairwise_forward_and_grads(golden_dir, kind, dt, tol):
    g = _load(golden_dir, f"pairwise_{kind}.npz")
    user, item, bias = (g[k].astype(dt) for k in ("user", "item", "bias"))
    ids = (g["uid"], g["pid"], g["nid"])
    if kind == "bpr":
        loss, l2 = O.bpr_forward(user, item, bias, *ids)
        gr = O.bpr_grads(user, item, bias, *ids)
    else:
        loss, l2 = O.ucml_forward(user, item, bias, *ids, margin=0.5)
        gr = O.ucml_grads(user, item, bias, *ids, margin=0.5)
        # the fixture must exercise both hinge branches
        assert 0 < np.count_nonzero(gr["g"]) < len(gr["g"])
    assert loss.dtype == dt
    np.testing.assert_allclose(loss, g["loss"], rtol=tol * 10, atol=tol)
    np.testing.assert_allclose(l2, g["l2"], rtol=tol * 10, atol=tol)
    for name, var in (("user", user), ("item", item), ("bias", bias)):
        idx, val = gr[name]
        np.testing.assert_allclose(_dense(len(var), idx, val.reshape(len(idx), -1)), g[f"g_{name}"],
                                   atol=tol * 20 if kind == "ucml" else tol, rtol=tol)


def test_dedup_matches_dense_scatter(golden_dir):
    g = _load(golden_dir, "pairwise_bpr.npz")
    gr = O.bpr_grads(g["user"], g["item"], g["bias"], g["uid"], g["pid"], g["nid"])
    idx, val = gr["item"]
    uniq, summed = O.dedup(idx, val)
    assert len(set(uniq.tolist())) == len(uniq) < len(idx)  # fixture has duplicates
    # tf.unique order = first occurrence
    seen = []
    for v in idx.tolist():
        if v not in seen:
            seen.append(v)
    assert uniq.tolist() == seen
    np.testing.assert_allclose(_dense(len(g["item"]), uniq, summed), g["g_item"], atol=1e-12)


def test_censor_and_inference(golden_dir):
    g = _load(golden_dir, "pairwise_ucml.npz")
    user, item = g["user"].copy(), g["item"].copy()
    O.ucml_censor_vec(user, item, g["uid"], g["pid"], g["nid"])
    np.testing.assert_allclose(user, g["user_censored"], atol=1e-12)
    np.testing.assert_allclose(item, g["item_censored"], atol=1e-12)
    untouched = np.setdiff1d(np.arange(len(user)), g["uid"])
    assert len(untouched) and np.array_equal(user[untouched], g["user"][untouched])
    np.testing.assert_allclose(O.ucml_inference(g["user"], g["item"], g["bias"], g["uid"][:5]),
                               g["inference"], atol=1e-12)
    b = _load(golden_dir, "pairwise_bpr.npz")
    np.testing.assert_allclose(O.dot_inference(b["user"], b["item"], b["bias"], b["uid"][:5]),
                               b["inference"], atol=1e-12)


@pytest.mark.parametrize("kind", ["gmf", "wrmf"])
@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-12), (np.float32, 2e-5)])
def test_pointwise_forward_and_grads(golden_dir, kind, dt, tol):
    g = _load(golden_dir, f"pointwise_{kind}.npz")
    user, item, bias = (g[k].astype(dt) for k in ("user", "item", "bias"))
    uid, iid, label = g["uid"], g["iid"], g["label"]
    if kind == "gmf":
        w = g["w"].astype(dt)
        loss, l2 = O.gmf_forward(user, item, bias, w, uid, iid, label)
        gr = O.gmf_grads(user, item, bias, w, uid, iid, label)
        np.testing.assert_allclose(gr["w"], g["g_w"], atol=tol, rtol=tol)
        np.testing.assert_allclose(O.gmf_inference(user, item, bias, w, uid[:5]), g["inference"], atol=tol)
    else:
        a, b = float(g["a"]), float(g["b"])
        loss, l2 = O.wrmf_forward(user, item, bias, uid, iid, label, a, b)
        gr = O.wrmf_grads(user, item, bias, uid, iid, label, a, b)
        np.testing.assert_allclose(O.dot_inference(user, item, bias, uid[:5]), g["inference"], atol=tol)
    np.testing.assert_allclose(loss, g["loss"], rtol=tol * 10, atol=tol)
    np.testing.assert_allclose(l2, g["l2"], rtol=tol * 10, atol=tol)
    for name, var in (("user", user), ("item", item), ("bias", bias)):
        idx, val = gr[name]
        np.testing.assert_allclose(_dense(len(var), idx, val.reshape(len(idx), -1)), g[f"g_{name}"],
                                   atol=tol * 10, rtol=tol)


def test_interaction_bug_compatible(golden_dir):
    """the reference's dot interaction is identically zero (diag only with self)."""
    g = _load(golden_dir, "interaction.npz")
    feats = [g[f"in{k}"] for k in range(5)]
    for si in (False, True):
        ref = g[f"out_self{int(si)}"]
        np.testing.assert_allclose(O.second_order_interaction(feats, si, "reference"), ref, atol=1e-12)
    assert not g["out_self0"].any()
    assert np.count_nonzero(g["out_self1"]) == 5 * 6  # only the F squared norms per sample
    fixed = O.second_order_interaction(feats, False, "dlrm")
    Z = np.stack(feats, 1)
    assert fixed.shape == (6, 10) and np.allclose(fixed[:, 0], (Z[:, 1] * Z[:, 0]).sum(-1))


@pytest.mark.parametrize("tag,kw", [("mse", {}), ("bce_self", dict(loss_func="bce", self_interaction=True)),
                                    ("clip", dict(loss_threshold=0.45)), ("bce", dict(loss_func="bce"))])
def test_dlrm_forward_backward(golden_dir, tag, kw):
    g = _load(golden_dir, f"dlrm_{tag}.npz")
    nv = int(g["n_vars"])
    var = [g[f"var{k}"] for k in range(nv)]
    ref = [g[f"grad{k}"] for k in range(nv)]
    # creation order (dlrm.py:32-37, keras builds Dense lazily at first call): 3 tables, then
    # bottom kernels/biases in call order, then top.
    tabs = var[:3]
    rest = var[3:]
    bot_w, bot_b = [rest[0], rest[2]], [rest[1], rest[3]]
    top_w, top_b = [rest[4], rest[6], rest[8]], [rest[5], rest[7], rest[9]]
    loss_func = kw.pop("loss_func", "mse")
    si = kw.pop("self_interaction", False)
    thr = kw.pop("loss_threshold", 0.0)
    cache = O.dlrm_forward(tabs, bot_w, bot_b, top_w, top_b, g["dense"], g["sparse"],
                           self_interaction=si, loss_threshold=thr)
    np.testing.assert_allclose(cache["pred"], g["pred"], atol=1e-12)
    loss, dpred = O.dlrm_loss(cache["pred"], g["label"], loss_func)
    np.testing.assert_allclose(loss, g["loss"], atol=1e-12)
    gr = O.dlrm_backward(cache, tabs, bot_w, top_w, g["dense"], g["sparse"], dpred, self_interaction=si)
    for k in range(3):
        np.testing.assert_allclose(_dense(len(tabs[k]), g["sparse"][:, k], gr["emb"][k]), ref[k], atol=1e-12)
        if not si:
            assert not ref[k].any()  # Q1: sparse tables receive exactly-zero gradients
    got = [gr["bot_w"][0], gr["bot_b"][0], gr["bot_w"][1], gr["bot_b"][1]]
    for l in range(3):
        got += [gr["top_w"][l], gr["top_b"][l]]
    for a, b in zip(got, ref[3:]):
        np.testing.assert_allclose(a, b, atol=1e-12)


def test_dlrm_fixed_interaction_gradcheck():
    """mode='dlrm' has no reference golden (the reference never computes it): check the
    closed-form backward against finite differences."""
    rng = np.random.default_rng(0)
    feats = [rng.standard_normal((3, 4)) for _ in range(4)]
    dout = rng.standard_normal((3, 6))
    dZ = O.second_order_interaction_bwd(feats, dout, False, "dlrm")
    eps = 1e-6
    for k in range(4):
        for idx in [(0, 1), (2, 3)]:
            fp = [f.copy() for f in feats]
            fm = [f.copy() for f in feats]
            fp[k][idx] += eps
            fm[k][idx] -= eps
            num = ((O.second_order_interaction(fp, False, "dlrm") - O.second_order_interaction(fm, False, "dlrm"))
                   * dout).sum() / (2 * eps)
            assert abs(num - dZ[idx[0], k, idx[1]]) < 1e-6


def test_metrics(golden_dir):
    g = _load(golden_dir, "metrics.npz")
    np.testing.assert_allclose(O.auc(g["pos"], g["pred"], g["excl"]), g["auc"], atol=1e-6)
    np.testing.assert_allclose(O.ndcg(g["pos"], g["pred"], g["excl"], (5, 20)), g["ndcg"], atol=1e-5)
    np.testing.assert_allclose(O.recall(g["pos"], g["pred"], g["excl"], (5, 20)), g["recall"], atol=1e-6)


def test_optimizer_known_answers():
    """Hand-computed known answers for the [TF-mem] optimizer formulas."""
    var = np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0]])
    idx = np.array([2, 0, 2])
    val = np.array([[1.0, 1.0], [0.5, 0.5], [2.0, 2.0]])
    v = var.copy()
    O.sgd_sparse(v, idx, val, lr=0.1)
    np.testing.assert_allclose(v, [[0.95, 1.95], [3, 4], [4.7, 5.7]])
    v, acc = var.copy(), np.full_like(var, 0.1)
    O.adagrad_sparse(v, acc, idx, val, lr=0.1, eps=0.0)
    np.testing.assert_allclose(acc, [[0.35, 0.35], [0.1, 0.1], [9.1, 9.1]])
    np.testing.assert_allclose(v[2], var[2] - 0.1 * 3 / np.sqrt(9.1))
    np.testing.assert_allclose(v[1], var[1])
    v, m, s = var.copy(), np.zeros_like(var), np.zeros_like(var)
    O.adam_dense_on_sparse(v, m, s, idx, val, step=1, lr=0.001, eps=1e-30)
    # first step, no eps: every touched row moves by exactly lr*sign(g); untouched rows by 0
    np.testing.assert_allclose(v, var - 0.001 * np.array([[1, 1], [0, 0], [1, 1]]), atol=1e-12)
    # second step with NO gradient on row 0: dense Adam still moves it (Q5), lazy Adam does not
    v2, m2, s2 = v.copy(), m.copy(), s.copy()
    O.adam_dense_on_sparse(v2, m2, s2, np.array([2]), np.array([[1.0, 1.0]]), step=2, eps=1e-30)
    assert np.all(v2[0] < v[0]) and np.all(m2[0] == 0.9 * m[0])
    v3, m3, s3 = v.copy(), m.copy(), s.copy()
    O.adam_lazy(v3, m3, s3, np.array([2]), np.array([[1.0, 1.0]]), step=2, eps=1e-30)
    assert np.array_equal(v3[0], v[0]) and np.array_equal(m3[0], m[0])


def test_train_step_uses_prestep_tables(golden_dir):
    """Duplicated rows: the step must gather everything before updating anything."""
    g = _load(golden_dir, "pairwise_bpr.npz")
    user, item, bias = g["user"].copy(), g["item"].copy(), g["bias"].copy()
    state = {k: (np.full_like(v, 0.1), None) for k, v in (("user", user), ("item", item), ("bias", bias))}
    loss, l2 = O.pairwise_train_step("bpr", user, item, bias, g["uid"], g["pid"], g["nid"],
                                     O.OPT_ADAGRAD, state, 1, lr=0.05)
    np.testing.assert_allclose(loss, g["loss"], atol=1e-12)
    acc_expect = 0.1 + g["g_item"] ** 2
    touched = np.union1d(g["pid"], g["nid"])
    np.testing.assert_allclose(state["item"][0][touched], acc_expect[touched], atol=1e-12)
    np.testing.assert_allclose(item[touched], g["item"][touched] - 0.05 * g["g_item"][touched]
                               / (np.sqrt(acc_expect[touched]) + 1e-7), atol=1e-12)
