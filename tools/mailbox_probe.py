"""Per-phase CUDA-event timing of the mailbox exchange step (run under torchrun, N >= 2):
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/mailbox_probe.py"""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
from openrec_b200 import _lib, native as N
from openrec_b200.sharded import MailboxShardedPairwise

eng = N.engine(torch.device("cuda", rank))
U, I, D, B = 1_000_000, 12_500_000 * world, 128, 65536
m = MailboxShardedPairwise(eng, rank, world, U, I, D, B, kind=0, opt_kind=N.ORX_OPT_ADAGRAD, lr=0.05, seed=1, barrier="flag")
g = torch.Generator(device="cpu").manual_seed(100 + rank)
ids = [tuple(torch.randint(0, n, (B,), generator=g, dtype=torch.int32).cuda() for n in (U, I, I)) for _ in range(8)]
vp = lambda t: C.c_void_p(t.data_ptr())
names = ["bucket", "push_ids", "barA", "gather_push", "barB", "grad_push", "barC", "apply"]
acc = [0.0] * len(names)
K = 40
for it in range(K + 5):
    uid, pid, nid = ids[it % 8]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    x, st = C.byref(m._x), eng.stream()
    m.iterations += 1
    ev[0].record()
    cat = torch.cat([uid, pid, nid])
    counts, send_local, slot = eng.owner_bucket_combined(cat, B, U, world)
    ev[1].record()
    _lib.check(eng.lib.orx_xchg_push_ids(eng.h, x, vp(counts), vp(send_local), 3 * B, st), "push")
    ev[2].record()
    m._barrier(); ev[3].record()
    _lib.check(eng.lib.orx_xchg_gather_push(eng.h, x, vp(m.table), m.table.shape[0], m.gin_rows, vp(m._req), vp(m._n_dev), None, st), "gp")
    ev[4].record()
    m._barrier(); ev[5].record()
    out4 = torch.zeros(4, device="cuda")
    _lib.check(eng.lib.orx_xchg_grad_push(eng.h, 0, x, vp(counts), vp(slot), B, D, 0.5, 1.0, 1.0, 1.0 / (B * world), vp(out4), st), "grad")
    ev[6].record()
    m._barrier(); ev[7].record()
    o = eng.make_opt(m.opt_kind, m.lr, m.eps, m.b1, m.b2, m.iterations)
    tab = eng.make_table(m.table, *m.slots)
    _lib.check(eng.lib.orx_sparse_apply_devn(eng.h, C.byref(tab), vp(m._req), C.c_void_p(m._bufs[3].ptr), m.W, m.gin_rows, vp(m._n_dev), C.byref(o), st), "apply")
    ev[8].record()
    torch.cuda.synchronize()
    if it >= 5:
        for k in range(len(names)):
            acc[k] += ev[k].elapsed_time(ev[k + 1])
if rank == 0:
    print("per-phase us (rank 0, per-step sync so no CPU run-ahead):", {n: round(a / K * 1e3, 1) for n, a in zip(names, acc)},
          "sum", round(sum(acc) / K * 1e3, 1))
# fused single-call step, free running
m.fused_call = True
for it in range(5):
    m.step(*ids[it % 8], reduce_loss=False)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(100):
    m.step(*ids[it % 8], reduce_loss=False)
e1.record(); torch.cuda.synchronize()
m.check()
if rank == 0:
    print("fused-call step: %.1f us" % (e0.elapsed_time(e1) * 10))
m.close()
dist.destroy_process_group()
