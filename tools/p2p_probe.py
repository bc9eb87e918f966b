"""Bandwidth through the CUDA-IPC peer mapping used by sharded_peer.py: bulk copy vs random 512 B rows (2 GPUs)."""
import ctypes as C, os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NCCL_DEBUG"] = "WARN"
from openrec_b200 import native as N, _lib
from openrec_b200.sharded_peer import _PeerBuf
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
eng = N.engine(dev)
rows, D = int(os.environ.get('P2P_ROWS', '4000000')), 128
buf = _PeerBuf(eng, (rows, D), torch.float32)
hs = [None] * world; dist.all_gather_object(hs, buf.handle)
other = (rank + 1) % world
p = C.c_void_p(); _lib.check(eng.lib.orx_peer_open(eng.h, hs[other], C.byref(p)))
class V:  # view of the peer allocation
    __cuda_array_interface__ = {"shape": (rows, D), "typestr": "<f4", "data": (p.value, False), "version": 2, "strides": None}
peer = torch.as_tensor(V(), device=dev)
local = torch.empty(rows, D, device=dev)
if rank == 0:
    print("can access peer:", torch.cuda.can_device_access_peer(0, 1))
def t(fn, nbytes, name, n=5):
    fn(); torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    if rank == 0: print(f"{name:44s} {nbytes*n/(e0.elapsed_time(e1)*1e-3)/1e9:8.1f} GB/s")
    dist.barrier()
sub = 400_000
t(lambda: local[:sub].copy_(peer[:sub]), sub * D * 4, "bulk peer READ  (copy_ 205 MB)")
t(lambda: peer[:sub].copy_(local[:sub]), sub * D * 4, "bulk peer WRITE (copy_ 205 MB)")
idx = torch.randint(0, rows, (196608,), device=dev)
t(lambda: torch.index_select(peer, 0, idx), 196608 * D * 4, "random 512B row peer READ (index_select)")
src = torch.randn(196608, D, device=dev)
t(lambda: peer.index_copy_(0, idx, src), 196608 * D * 4, "random 512B row peer WRITE (index_copy_)")
idx32 = idx.to(torch.int32)
t(lambda: eng.gather(peer, idx32), 196608 * D * 4, "random 512B row peer READ (orx_gather)")
t(lambda: eng.gather(local, idx32), 196608 * D * 4, "random 512B row LOCAL read (orx_gather)")
small = torch.zeros(rows, device=dev)
class V1:
    __cuda_array_interface__ = {"shape": (rows * D,), "typestr": "<f4", "data": (p.value, False), "version": 2, "strides": None}
peer1 = torch.as_tensor(V1(), device=dev)
vals = torch.randn(393216, device=dev); i2 = torch.randint(0, rows * D, (393216,), device=dev)
t(lambda: peer1.index_copy_(0, i2, vals), 393216 * 4, "random 4-byte peer WRITES (393k)")
torch.cuda.synchronize(); dist.barrier()
eng.lib.orx_peer_close(eng.h, p); dist.barrier(); buf.free(); dist.destroy_process_group()
