"""Per-phase timing of PeerShardedPairwise.step (torchrun, >= 2 GPUs)."""
import ctypes as C, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NCCL_DEBUG"] = "WARN"
from openrec_b200 import native as N, _lib
from openrec_b200.sharded_peer import PeerShardedPairwise

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
eng = N.engine(dev)
U, I, D, B = 1_000_000, 12_500_000 * world, 128, 65536
m = PeerShardedPairwise(eng, rank, world, U, I, D, B, kind=0, opt_kind=1, lr=0.05, seed=1)
g = torch.Generator().manual_seed(100 + rank)
ids = [tuple(torch.randint(0, n, (B,), generator=g, dtype=torch.int32).to(dev) for n in (U, I, I)) for _ in range(8)]
for i in range(5):
    m.step(*ids[i % 8], reduce_loss=False)
torch.cuda.synchronize(); dist.barrier()
acc, names = {}, []
def tick(name, t0):
    torch.cuda.synchronize()
    t = time.perf_counter()
    acc[name] = acc.get(name, 0.0) + (t - t0)
    if name not in names: names.append(name)
    return time.perf_counter()
n_it = 30
for it in range(n_it):
    uid, pid, nid = ids[it % 8]
    dist.barrier(); torch.cuda.synchronize()
    t = time.perf_counter()
    out4 = torch.zeros(4, device=dev)
    m._barrier(); t = tick("barrier 1", t)
    _lib.check(eng.lib.orx_peer_pairwise_push(eng.h, 0, C.byref(m._peer), C.c_void_p(uid.data_ptr()), C.c_void_p(pid.data_ptr()),
               C.c_void_p(nid.data_ptr()), B, C.c_void_p(m._pos.data_ptr()), 0.5, 1.0, 1.0, 1.0 / (B * world),
               C.c_void_p(out4.data_ptr()), eng.stream())); t = tick("push (bucket + peer kernel)", t)
    m._barrier(); t = tick("barrier 2", t)
    o = eng.make_opt(1, 0.05, 1e-7, 0.9, 0.999, it + 1)
    _lib.check(eng.lib.orx_peer_apply(eng.h, C.byref(m._tab(m.emb, m.emb_slots)), C.byref(m._tab(m.bias.reshape(-1, 1), m.bias_slots)),
               C.c_void_p(m._in_ids.ptr), C.c_void_p(m._in_emb.ptr), C.c_void_p(m._in_bias.ptr), C.c_void_p(m._in_cnt.ptr),
               world, m.cap, C.byref(o), eng.stream())); t = tick("apply (index + apply + tail)", t)
if rank == 0:
    for k in names: print(f"{k:32s} {acc[k]/n_it*1e6:9.1f} us")
    print(f"{'sum':32s} {sum(acc.values())/n_it*1e6:9.1f} us")
torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
for it in range(100): m.step(*ids[it % 8], reduce_loss=False)
torch.cuda.synchronize()
if rank == 0: print(f"free-running step {(time.perf_counter()-t0)/100*1e6:.1f} us")
m.close(); dist.destroy_process_group()
