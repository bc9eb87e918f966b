"""Worker of tests/test_sharded_gloo.py (CPU, gloo, oracle-backed engine) and tests/test_gpu_sharded.py (one process per
GPU, the real kernels): one rank of a world_size-R job running three steps of the row-sharded pairwise step."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    out_path, kind, opt_kind = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    mode = sys.argv[4] if len(sys.argv) > 4 else "cpu"
    on_gpu = mode != "cpu"
    from openrec_b200.sharded import HomeRoutedPairwise, ShardedPairwise
    if on_gpu:   # the real kernels, one process per GPU
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from openrec_b200 import native
        engine = native.engine(torch.device("cuda", rank))
    else:        # host logic only: oracle-backed stand-in over gloo
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from fake_engine import FakeEngine
        engine = FakeEngine()
    dev = engine.device
    rng = np.random.default_rng(99)                       # same global problem on every rank
    U, I, D, B = (61, 83, 16, 40) if not on_gpu else (1501, 2003, 128, 1024)
    sc = 0.05 if kind == 0 else 0.4
    user, item, bias = (rng.uniform(-sc, sc, s).astype(np.float32) for s in ((U, D), (I, D), (I, 1)))
    if mode in ("home", "home_next"):   # peer-store mailboxes, home-routed (csrc/orx_shard.cu)
        m = HomeRoutedPairwise(engine, rank, world, U, I, D, B, kind=kind, opt_kind=opt_kind, lr=0.05, init=False)
    else:                # NCCL / gloo all-to-all form
        m = ShardedPairwise(engine, rank, world, U, I, D, kind=kind, opt_kind=opt_kind, lr=0.05, init=False)
    m.load_global(user, item, bias)
    losses = []
    steps = 5 if mode == "home_next" else 3
    batches = []
    for step in range(steps):
        ids = [rng.integers(0, n, B * world).astype(np.int32) for n in (U, I, I)]   # global batch
        batches.append([torch.from_numpy(a[rank * B:(rank + 1) * B].copy()).to(dev) for a in ids])
    for step in range(steps):
        if mode == "home_next":   # the next batch announced a step ahead: its route / request ride in this step's apply launch
            nxt = batches[step + 1] if step + 1 < steps and step != 2 else None
            losses.append(m.step(*batches[step], next_ids=nxt).cpu().numpy().copy())
        else:
            losses.append(m.step(*batches[step]).cpu().numpy().copy())
    full = [t.cpu().numpy() for t in m.gather_global()]
    if rank == 0:
        np.savez(out_path, user=full[0], item=full[1], bias=full[2], losses=np.stack(losses))
    if mode in ("cpu", "home", "home_next") and opt_kind in (0, 1) and kind == 0:
        # the touched-rows check that bench.py --gpus N --check runs at the full table shape (tests/shard_check.py):
        # one more step, verified on rank 0 against the oracle from the shards' own snapshots
        import shard_check
        access = shard_check.HomeRoutedAccess(m) if mode != "cpu" else shard_check.CombinedAccess(m)
        verdict = shard_check.run(m, access, rank, world, U, I, D, B, kind=kind, opt_kind=opt_kind, lr=0.05,
                                  atol=1e-5 if kind == 0 else 2e-4)
        if rank == 0:
            assert verdict["passed"] and verdict["global_batch"] == B * world
    if hasattr(m, "check"):
        m.check()
    if hasattr(m, "close"):
        m.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
