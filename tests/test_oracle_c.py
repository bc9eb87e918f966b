"""CPU: the C/OpenMP port (cpu_baseline arm) agrees with the numpy oracle."""
import numpy as np
import pytest

from oracle import c_port
from oracle import openrec_oracle as O


@pytest.mark.parametrize("kind", ["bpr", "ucml"])
@pytest.mark.parametrize("opt", [0, 1])
def test_c_port_matches_numpy_oracle(kind, opt):
    rng = np.random.default_rng(7)
    U, I, D, B = 200, 300, 64, 1000
    sc = 0.05 if kind == "bpr" else 0.3
    user = rng.uniform(-sc, sc, (U, D)).astype(np.float32)
    item = rng.uniform(-sc, sc, (I, D)).astype(np.float32)
    bias = rng.uniform(-sc, sc, (I, 1)).astype(np.float32)
    uid, pid, nid = (rng.integers(0, n, B).astype(np.int32) for n in (U, I, I))
    ref = [a.astype(np.float64) for a in (user, item, bias)]
    st = {k: (np.full_like(v, 0.1), None) for k, v in zip(("user", "item", "bias"), ref)}
    loss, l2 = O.pairwise_train_step(kind, *ref, uid, pid, nid, opt, st, 1, 0.05, margin=0.5)
    acc = [np.full_like(a, 0.1) for a in (user, item, bias)]
    closs, cl2 = c_port.pairwise_step(kind, user, acc[0], item, acc[1], bias, acc[2], uid, pid, nid, opt, 0.05)
    np.testing.assert_allclose(closs, loss, rtol=2e-5)
    np.testing.assert_allclose(cl2, l2, rtol=2e-5)
    for got, want in zip((user, item, bias), ref):
        np.testing.assert_allclose(got, want, atol=2e-5)
    if opt == 1:
        np.testing.assert_allclose(acc[1], st["item"][0], atol=2e-5, rtol=1e-5)
    assert c_port.num_threads() >= 1
