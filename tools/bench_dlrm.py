"""DLRM Criteo-shape step (BASELINE configs[3]): 26 sparse features x 1M vocab x dim 128, 13 dense, batch 32768,
bottom MLP 13-512-256-128, top MLP 479-1024-1024-512-256-1 (sizes are this repo's choice), Adagrad,
interaction mode 'dlrm' (the reference's own interaction is identically zero).
    PYTHONPATH=compat:. python tools/bench_dlrm.py [--steps 20] [--simt]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "compat"), ROOT]
ap = argparse.ArgumentParser(); ap.add_argument("--steps", type=int, default=20); ap.add_argument("--simt", action="store_true")
ap.add_argument("--vocab", type=int, default=1_000_000); ap.add_argument("--batch", type=int, default=32768)
a = ap.parse_args()
if a.simt: os.environ["ORX_MLP_SIMT"] = "1"
import numpy as np, torch, tensorflow as tf
from openrec.tf2.recommenders import DLRM
rng = np.random.default_rng(0)
B, T, D = a.batch, 26, 128
model = DLRM(m_spa=D, ln_emb=[a.vocab] * T, ln_bot=[512, 256, D], ln_top=[1024, 1024, 512, 256, 1], interaction_mode="dlrm")
opt = tf.keras.optimizers.Adagrad(learning_rate=0.01)
batches = [(tf.constant(np.log1p(rng.integers(0, 100, (B, 13))).astype(np.float32)),
            tf.constant(rng.integers(0, a.vocab, (B, T)).astype(np.int32)),
            tf.constant((rng.random(B) < 0.25).astype(np.float32))) for _ in range(4)]
def step(i):
    d, s, y = batches[i % 4]
    with tf.GradientTape() as tape:
        loss = model(d, s, y)
    g = tape.gradient(loss, model.trainable_variables)
    opt.apply_gradients(zip(g, model.trainable_variables))
    return loss
for i in range(3): last = step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(a.steps): last = step(i)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
flops = 3 * 2 * B * (13*512 + 512*256 + 256*128 + 479*1024 + 1024*1024 + 1024*512 + 512*256 + 256)
print(json.dumps({"metric": "dlrm_samples_per_sec", "value": B / (ms * 1e-3), "ms_per_step": ms, "batch": B,
                  "mlp": "simt fp32" if a.simt else "tcgen05 3xTF32", "mlp_tflops_fp32_equiv": flops / (ms * 1e-3) / 1e12,
                  "loss": float(last), "config": "26 x %d x 128 tables, Adagrad, interaction_mode=dlrm" % a.vocab}))
