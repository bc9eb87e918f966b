"""CPU host logic of openrec_b200.tf2.checkpoint (row N4): save -> load into a fresh model/optimizer resumes the run
exactly.  Runs in a subprocess because tests/fake_engine.install() re-routes the engine process-wide."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path[:0] = [{compat!r}, {root!r}, {tests!r}]
import numpy as np, torch
import fake_engine
fake_engine.install()
import tensorflow as tf
from openrec.tf2.recommenders import BPR
from openrec_b200.tf2 import checkpoint

def step(model, opt, ids):
    with tf.GradientTape() as tape:
        out = model(*ids)
    g = tape.gradient(out, model.trainable_variables)
    opt.apply_gradients(zip(g, model.trainable_variables))
    return float(out[0]), float(out[1])

rng = np.random.default_rng(3)
U, I, D, B = 40, 60, 8, 64
mk = lambda: tuple(torch.from_numpy(rng.integers(0, n, B).astype(np.int32)) for n in (U, I, I))
a, b = mk(), mk()
m1, o1 = BPR(D, D, U, I), tf.keras.optimizers.Adagrad(learning_rate=0.05)
step(m1, o1, a)
checkpoint.save({path!r}, m1, o1)
m2, o2 = BPR(D, D, U, I), tf.keras.optimizers.Adagrad(learning_rate=0.05)
checkpoint.load({path!r}, m2, o2)
assert o2.iterations == o1.iterations == 1
for v1, v2 in zip(m1.variables, m2.variables):
    assert np.array_equal(v1.numpy(), v2.numpy())
l1, l2 = step(m1, o1, b), step(m2, o2, b)
assert l1 == l2, (l1, l2)
for v1, v2 in zip(m1.variables, m2.variables):
    assert np.array_equal(v1.numpy(), v2.numpy())            # Adagrad accumulators were restored too
try:
    checkpoint.load({path!r}, BPR(D + 1, D + 1, U, I))
    raise SystemExit("shape mismatch not detected")
except ValueError:
    pass
print("checkpoint ok")
"""


def test_checkpoint_roundtrip_cpu(tmp_path):
    code = SCRIPT.format(compat=os.path.join(ROOT, "compat"), root=ROOT, tests=os.path.join(ROOT, "tests"),
                         path=str(tmp_path / "ck"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "checkpoint ok" in r.stdout, r.stdout + r.stderr
