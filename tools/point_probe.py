"""GMF pointwise step at the BASELINE configs[1] table shape (1M x 1M, D=128, B=65536, Adagrad), a few steps: the ncu
target for k_point_step (tools/gpu_ncu.sh), and a CUDA-event time of the step.   python tools/point_probe.py [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from openrec_b200 import native as N

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
dev = torch.device("cuda", 0)
eng = N.engine(dev)
U = I = 1_000_000
D, B = 128, 65536
mk = lambda r, c, k: (lambda t: (eng.fill_uniform(t, -0.05, 0.05, k), t)[1])(torch.empty(r, c, device=dev))
user, item, bias, w = mk(U, D, 1), mk(I, D, 2), mk(I, 1, 3), torch.ones(1, D, device=dev)
acc = [torch.full_like(t, 0.1) for t in (user, item, bias, w)]
tabs = [N.table(t, a) for t, a in zip((user, item, bias, w), acc)]
g = torch.Generator(device="cpu").manual_seed(7)
ids = [(torch.randint(0, U, (B,), generator=g, dtype=torch.int32).to(dev), torch.randint(0, I, (B,), generator=g, dtype=torch.int32).to(dev),
        (torch.rand(B, generator=g) < 0.5).float().to(dev)) for _ in range(8)]
out4 = torch.zeros(4, device=dev)


def step(k):
    u, i, y = ids[k % 8]
    eng.pointwise_step(N.ORX_POINT_GMF, *tabs, u, i, y, N.opt(N.ORX_OPT_ADAGRAD, 0.05, step=k + 1), out4)


for k in range(5):
    step(k)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(steps):
    step(k)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
alg = B * (8 + 4 * (2 * D + 1) * 4)     # ids + label, rows and accumulators read and written once
print({"gmf_ms_per_step": round(ms, 4), "samples_per_s": round(B / ms * 1e3), "algorithmic_GBs_step_level": round(alg / ms / 1e6, 1),
       "loss": out4[:2].tolist()})
