"""Single-GPU profile of the sharded step's kernels: R virtual ranks on one device at the bench's batch size (every "peer"
store is local, so the numbers are the HBM-only cost of each launch).  Also the ncu target for k_sh_* (one GPU).
    python tools/shard_loopback_probe.py [R] [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from openrec_b200.sharded import LoopbackGroup

R = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
U, I, D, B = 1_000_000, 4_000_000 * R, 128, 65_536
g = LoopbackGroup(R, U, I, D, B, kind=0, opt_kind=1, lr=0.05, seed=1)
gen = torch.Generator(device="cpu").manual_seed(7)
ids = [[tuple(torch.randint(0, n, (B,), generator=gen, dtype=torch.int32).cuda() for n in (U, I, I)) for _ in range(R)]
       for _ in range(4)]
names = ["route", "request", "serve", "compute", "apply", "tail"]
for i in range(3):
    g.step(ids[i % 4])
torch.cuda.synchronize()
acc = [0.0] * 6
for i in range(steps):
    for m in g.ranks:
        m.iterations += 1
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    ev[0].record()
    for ph in range(6):
        for r, m in enumerate(g.ranks):
            m._call(*ids[i % 4][r], 1.0, 1.0, ph, ph)
        ev[ph + 1].record()
    torch.cuda.synchronize()
    for ph in range(6):
        acc[ph] += ev[ph].elapsed_time(ev[ph + 1]) * 1e3
g.check()
print(json.dumps({"virtual_ranks": R, "per_rank_us": {n: round(a / steps / R, 1) for n, a in zip(names, acc)},
                  "sum_us_per_rank": round(sum(acc) / steps / R, 1)}))
g.close()
