"""Per-launch timing of HomeRoutedPairwise.step (torchrun, >= 2 GPUs): CUDA events between the launches inside the
one C call of a step (orx_profile_*; a launch that waits for a peer's flag includes that wait), once with the next batch
announced a step ahead (route / request ride in the apply launch) and once without.
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
cnt = [0]


def run(n, announce):
    for _ in range(n):
        k = cnt[0]
        cnt[0] += 1
        m.step(*ids[k % 8], next_ids=ids[(k + 1) % 8] if announce else None)


def finish():
    if m._announced is not None:
        m.step(*m._announced)
        cnt[0] += 1


out = {"rank": rank, "world": world}
for announce in (True, False):
    run(5, announce)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(steps, announce)          # whole step, one call
    e1.record(); torch.cuda.synchronize()
    whole = e0.elapsed_time(e1) / steps * 1e3
    dist.barrier()
    # per launch, inside the ONE C call of a step: orx_profile_* records events between the launches of every 8th step
    eng.profile_enable(True)
    run(8 * 24, announce)
    torch.cuda.synchronize()
    ms, n = eng.profile_read(6)
    eng.profile_enable(False)
    finish()
    key = "announced" if announce else "plain"
    out[key] = {"whole_step_us": round(whole, 1), "phase_us": {nm: round(x * 1e3 / max(n, 1), 1) for nm, x in zip(names, ms)},
                "sum_us": round(sum(ms) * 1e3 / max(n, 1), 1)}
m.check()
gathered = [None] * world
dist.all_gather_object(gathered, out)
if rank == 0:
    for o in gathered:
        print(json.dumps(o))
m.close()
dist.destroy_process_group()
