"""Per-launch timing of HomeRoutedPairwise.step (torchrun, >= 2 GPUs): CUDA events between the six launches inside the
one C call of a step (orx_profile_*; a launch that waits for a peer's flag includes that wait).
    torchrun --nproc-per-node N tools/shard_probe.py [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist

import bench as B
from openrec_b200 import native as N
from openrec_b200.sharded import HomeRoutedPairwise

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
eng = N.engine(dev)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
U, I, D, Bsz = B.U, 12_500_000 * world, B.D, B.B
m = HomeRoutedPairwise(eng, rank, world, U, I, D, Bsz, kind=0, opt_kind=1, lr=B.LR, seed=1)
g = torch.Generator(device="cpu").manual_seed(100 + rank)
ids = [tuple(torch.randint(0, n, (Bsz,), generator=g, dtype=torch.int32).to(dev) for n in (U, I, I)) for _ in range(8)]
names = ["route", "request", "serve", "compute", "apply", "tail"]
for i in range(5):
    m.step(*ids[i % 8])
dist.barrier(); torch.cuda.synchronize()
# whole step, one call
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(steps):
    m.step(*ids[i % 8])
e1.record(); torch.cuda.synchronize()
whole = e0.elapsed_time(e1) / steps * 1e3
dist.barrier()
# per launch, inside the ONE C call of a step: orx_profile_* records events between the six launches of every 8th step
eng.profile_enable(True)
for i in range(8 * 24):
    m.step(*ids[i % 8])
torch.cuda.synchronize()
ms, n = eng.profile_read(6)
eng.profile_enable(False)
acc = [x * 1e3 * steps / max(n, 1) for x in ms]
m.check()
out = {"rank": rank, "world": world, "whole_step_us": round(whole, 1),
       "phase_us": {n: round(a / steps, 1) for n, a in zip(names, acc)}, "sum_us": round(sum(acc) / steps, 1)}
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    for o in gathered:
        print(json.dumps(o))
m.close()
dist.destroy_process_group()
