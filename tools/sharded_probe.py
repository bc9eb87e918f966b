"""Per-phase timing of ShardedPairwise.step (torchrun, >= 2 GPUs): where does the sharded step spend its time?"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NCCL_DEBUG"] = "WARN"
from openrec_b200 import native as N
from openrec_b200.sharded import ShardedPairwise, _a2a

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
eng = N.engine(dev)
U, I, D, B = 1_000_000, 12_500_000 * world, 128, 65536
m = ShardedPairwise(eng, rank, world, U, I, D, kind=0, opt_kind=1, lr=0.05, seed=1)
g = torch.Generator().manual_seed(100 + rank)
ids = [tuple(torch.randint(0, n, (B,), generator=g, dtype=torch.int32).to(dev) for n in (U, I, I)) for _ in range(8)]
for i in range(5):
    m.step(*ids[i % 8], reduce_loss=False)
torch.cuda.synchronize(); dist.barrier()
names, acc = [], {}
def tick(name, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t - t0)
    if name not in names: names.append(name)
    return time.perf_counter()
R, W = world, m.W
n_it = 30
for it in range(n_it):
    uid, pid, nid = ids[it % 8]
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    cat = torch.cat([uid, pid, nid]); counts, send_local, slot = eng.owner_bucket_combined(cat, B, m.U, R); t = tick("bucket", t)
    rcounts = torch.empty_like(counts); dist.all_to_all_single(rcounts, counts); t = tick("a2a counts", t)
    host = torch.stack([counts, rcounts]).cpu(); sc, rc = host[0].tolist(), host[1].tolist(); t = tick("host sync", t)
    req = torch.empty(sum(rc), dtype=torch.int32, device=dev); _a2a(req, send_local, rc, sc); t = tick("a2a ids", t)
    rows = eng.gather(m.table, req); t = tick("gather", t)
    got = torch.empty(3 * B, W, dtype=torch.float32, device=dev); _a2a(got, rows, sc, rc); t = tick("a2a rows", t)
    out4 = torch.zeros(4, device=dev); d_got = torch.empty_like(got)
    eng.pairwise_grad_rows(0, got, D, slot[:B], slot[B:2*B], slot[2*B:], 1.0/(B*R), d_got, out4, 0.5, 1.0, 1.0); t = tick("grad kernel", t)
    g_rows = torch.empty_like(rows); _a2a(g_rows, d_got, rc, sc); t = tick("a2a grads", t)
    o = eng.make_opt(1, 0.05, 1e-7, 0.9, 0.999, it + 1)
    eng.sparse_apply(eng.make_table(m.table, *m.slots), req, g_rows, o); t = tick("sparse_apply", t)
if rank == 0:
    tot = sum(acc.values())
    for k in names: print(f"{k:14s} {acc[k]/n_it*1e6:9.1f} us")
    print(f"{'sum':14s} {tot/n_it*1e6:9.1f} us  (each phase individually synchronised)")
# free-running
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for it in range(100): m.step(*ids[it % 8], reduce_loss=False)
torch.cuda.synchronize(); 
if rank == 0: print(f"free-running step {(time.perf_counter()-t0)/100*1e6:.1f} us")
dist.destroy_process_group()
