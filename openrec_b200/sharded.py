"""Row-sharded BPR / UCML step across the GPUs of one NVSwitch box (BASELINE configs[4]).

The reference is single-device; this is the scale-out of the same synchronous step:
row r of every table lives on rank ``r % R`` at local row ``r // R`` (optimizer slots alongside);
each rank owns B triplets of the global batch.  One step =

  1. orx_owner_bucket_combined: sort this rank's 3B lookups by owner                   (liborx)
  2. all-to-all            : lookup counts, then combined local-row ids, to the owners (NCCL over NVLink)
  3. orx_gather            : owners read the requested rows from their shard          (liborx)
  4. all-to-all            : rows back to the requesters
  5. orx_pairwise_grad_rows: score, loss and per-lookup gradient rows
                             (pre-step values everywhere: nothing has been written yet)
  6. all-to-all            : gradient rows to the owners
  7. orx_sparse_apply      : owners sum duplicates (across ALL ranks' lookups) and apply the
                             optimizer once per unique row                             (liborx)
  8. all-reduce            : (loss, l2_loss), only when the caller asks for the global value

``torch.distributed`` is plumbing (one process per GPU, NCCL; gloo in the CPU logic tests); every
arithmetic op is a liborx kernel reached through ``eng``.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


def _a2a(out, inp, out_splits, in_splits):
    dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits)


class ShardedPairwise:
    """BPR (kind 0) / UCML (kind 1) with row-sharded tables.  ``eng`` is a native.Engine (or the
    oracle-backed stand-in of tests/fake_engine.py on CPU).

    Local storage is ONE combined table ``[user rows | item rows, D+4]`` per rank (item bias in column D, three
    padding columns keep rows 16-byte aligned), with optimizer slots of the same shape: a lookup is then just
    (owner, combined local row), and a step needs four collectives -- counts, ids, rows, gradient rows."""

    PAD = 4

    def __init__(self, eng, rank, world, total_users, total_items, dim, *, kind=0, opt_kind=1, lr=0.05, eps=1e-7,
                 beta1=0.9, beta2=0.999, margin=0.5, seed=0, init=True):
        self.eng, self.rank, self.world = eng, rank, world
        self.U, self.I, self.D = total_users, total_items, dim
        self.W = dim + self.PAD
        self.kind, self.opt_kind, self.lr, self.eps, self.b1, self.b2, self.margin = kind, opt_kind, lr, eps, beta1, beta2, margin
        self.iterations = 0
        dev = eng.device
        self.ru = (total_users - rank + world - 1) // world     # rows r with r % world == rank
        self.ri = (total_items - rank + world - 1) // world
        self.table = torch.zeros(self.ru + self.ri, self.W, dtype=torch.float32, device=dev)
        if init:
            tmp = torch.empty(self.ru + self.ri, dim + 1, dtype=torch.float32, device=dev)
            eng.fill_uniform(tmp, -0.05, 0.05, seed * 1000003 + rank * 17)
            self.table[:, :dim + 1] = tmp
            self.table[:self.ru, dim] = 0.0                      # users have no bias column
            del tmp
        n_slots = {0: 0, 1: 1, 2: 2, 3: 2}[opt_kind]
        fill = 0.1 if opt_kind == 1 else 0.0
        self.slots = [torch.full_like(self.table, fill) for _ in range(n_slots)] + [None] * (2 - n_slots)
        self.launches_per_step = 3 + 1 + 2 + 3   # bucket(3) gather(1) grad+reduce(2) sparse_apply(3)
        self._loss_local = None

    def step(self, uid, pid, nid, c_loss=1.0, c_l2=1.0, reduce_loss=True):
        """uid/pid/nid: this rank's int32 GLOBAL ids on the device.  Returns a [2] device tensor: the global
        (loss, l2_loss) when ``reduce_loss`` (one extra all-reduce), else this rank's partial sums."""
        eng, R, D, W = self.eng, self.world, self.D, self.W
        B = uid.numel()
        dev = uid.device
        self.iterations += 1
        ids = torch.cat([uid, pid, nid])
        counts, send_local, slot = eng.owner_bucket_combined(ids, B, self.U, R)
        rcounts = torch.empty_like(counts)
        dist.all_to_all_single(rcounts, counts)
        host = torch.stack([counts, rcounts]).cpu()                    # the step's one host sync
        sc, rc = host[0].tolist(), host[1].tolist()
        req = torch.empty(sum(rc), dtype=torch.int32, device=dev)
        _a2a(req, send_local, rc, sc)
        rows = eng.gather(self.table, req)                             # owners read their shard
        got = torch.empty(3 * B, W, dtype=torch.float32, device=dev)
        _a2a(got, rows, sc, rc)
        out4 = torch.zeros(4, dtype=torch.float32, device=dev)
        d_got = torch.empty_like(got)
        eng.pairwise_grad_rows(self.kind, got, D, slot[:B], slot[B:2 * B], slot[2 * B:], 1.0 / (B * R), d_got, out4,
                               self.margin, c_loss, c_l2)
        g_rows = torch.empty_like(rows)
        _a2a(g_rows, d_got, rc, sc)
        o = eng.make_opt(self.opt_kind, self.lr, self.eps, self.b1, self.b2, self.iterations)
        eng.sparse_apply(eng.make_table(self.table, *self.slots), req, g_rows, o)   # dedup across ALL ranks' lookups
        out = out4[:2].clone()
        if reduce_loss:
            dist.all_reduce(out)
        return out

    # ---- helpers for tests: assemble / scatter the global tables
    def load_global(self, user, item, bias):
        r, R, D = self.rank, self.world, self.D
        self.table[:self.ru, :D] = torch.as_tensor(user[r::R], dtype=torch.float32)
        self.table[self.ru:, :D] = torch.as_tensor(item[r::R], dtype=torch.float32)
        self.table[self.ru:, D] = torch.as_tensor(bias[r::R], dtype=torch.float32).reshape(-1)

    def gather_global(self):
        """-> (user, item, bias) full tables on every rank (test helper; sizes must be small)."""
        D = self.D
        outs = []
        for t, total in ((self.table[:self.ru, :D], self.U), (self.table[self.ru:, :D], self.I),
                         (self.table[self.ru:, D:D + 1], self.I)):
            per = (total + self.world - 1) // self.world
            pad = torch.zeros(per, t.shape[1], dtype=t.dtype, device=t.device)
            pad[:t.shape[0]] = t
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad)
            outs.append(torch.stack(parts, 1).reshape(per * self.world, t.shape[1])[:total])   # row = local*R + rank
        return outs


class MailboxShardedPairwise(ShardedPairwise):
    """Same partitioning and arithmetic as ShardedPairwise, but the three exchanges are liborx kernels that STORE into
    the peers' mailboxes over NVLink (CUDA IPC mappings; csrc/orx_xchg.cu) -- no NCCL in the data path, no host sync.
    ``torch.distributed`` is used once to swap the IPC handles, and (``barrier="nccl"``) for the three tiny
    stream-ordered barriers of a step; ``barrier="flag"`` uses orx_xchg_barrier (flags in peer memory) instead.

    Buffer reuse across steps is ordered by the same barriers: idbox is rewritten after C(t) and read before B(t);
    got is rewritten after A(t+1) and read before C(t); gin is rewritten after B(t+1) and read by the owner's apply
    of step t, which precedes its arrival at A(t+1)."""

    def __init__(self, eng, rank, world, total_users, total_items, dim, batch, *, barrier=None, gin_rows=None, **kw):
        super().__init__(eng, rank, world, total_users, total_items, dim, **kw)
        from .sharded_peer import _PeerBuf
        if dim % 4:
            raise ValueError("the mailbox exchange needs dim % 4 == 0")
        self.B = batch
        self.cap = 3 * batch
        self.gin_rows = int(gin_rows or world * self.cap)        # worst case: every rank's lookups land on me
        self.barrier_kind = barrier or os.environ.get("ORX_XCHG_BARRIER", "flag")
        dev = eng.device
        W = self.W
        self._bufs = [_PeerBuf(eng, (world * self.cap,), torch.int32),      # idbox
                      _PeerBuf(eng, (world * 8,), torch.int32),             # meta
                      _PeerBuf(eng, (self.cap, W), torch.float32),          # got
                      _PeerBuf(eng, (self.gin_rows, W), torch.float32),     # gin
                      _PeerBuf(eng, (world + 1,), torch.int32)]             # flags
        for b in self._bufs:
            b.t.zero_()
        torch.cuda.synchronize()
        everyone = [None] * world
        dist.all_gather_object(everyone, [b.handle for b in self._bufs])
        self._opened = []
        ptrs = np.zeros((5, world), dtype=np.int64)
        for r in range(world):
            for k in range(5):
                if r == rank:
                    ptrs[k, r] = self._bufs[k].ptr
                else:
                    p = C.c_void_p()
                    _lib.check(eng.lib.orx_peer_open(eng.h, everyone[r][k], C.byref(p)), "orx_peer_open")
                    self._opened.append(p.value)
                    ptrs[k, r] = p.value
        self._ptrs = torch.from_numpy(ptrs).to(dev)
        base = self._ptrs.data_ptr()
        self._x = _lib.OrxXchg(world, rank, W, self.cap, *[base + 8 * world * k for k in range(5)])
        self._req = torch.empty(self.gin_rows, dtype=torch.int32, device=dev)
        self._n_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._flag = torch.zeros(1, dtype=torch.float32, device=dev)
        self._epoch = 0
        self._work = torch.zeros(world + 1 + ((self.cap + 1023) // 1024) * world, dtype=torch.int32, device=dev)
        self._slot = torch.empty(self.cap, dtype=torch.int32, device=dev)
        self._out = torch.zeros(16, 4, dtype=torch.float32, device=dev)     # ring of step outputs (loss, l2, -, -)
        self.fused_call = os.environ.get("ORX_XCHG_FUSED", "1") != "0" and self.barrier_kind == "flag"
        self.launches_per_step = 2 + 1 + 2 + 3 + 3 + 1   # hist, scatter+push, gather+push, grads+loss push, apply(3), 3 barriers, loss sum
        dist.barrier()

    def _barrier(self):
        if self.barrier_kind == "flag":
            self._epoch += 1
            _lib.check(self.eng.lib.orx_xchg_barrier(self.eng.h, C.byref(self._x), self._epoch, 5000, self.eng.stream()),
                       "orx_xchg_barrier")
        else:
            dist.all_reduce(self._flag)

    def step(self, uid, pid, nid, c_loss=1.0, c_l2=1.0, reduce_loss=True):
        eng, R, D, W = self.eng, self.world, self.D, self.W
        B = uid.numel()
        if B > self.B:
            raise ValueError("batch larger than the mailboxes this model was built for")
        self.iterations += 1
        x, st = C.byref(self._x), eng.stream()
        vp = lambda t: C.c_void_p(t.data_ptr())
        if self.fused_call:       # one C call: bucket + push, gather + push, grads + push, apply, 3 flag barriers
            out4 = self._out[self.iterations % 16]
            o = eng.make_opt(self.opt_kind, self.lr, self.eps, self.b1, self.b2, self.iterations)
            tab = eng.make_table(self.table, *self.slots)
            _lib.check(eng.lib.orx_xchg_step(eng.h, self.kind, x, C.byref(tab), vp(uid), vp(pid), vp(nid), B, self.U, D,
                                             C.c_void_p(self._bufs[3].ptr), self.gin_rows, vp(self._work),
                                             vp(self._slot), vp(self._req), self.margin, c_loss, c_l2, 1.0 / (B * R),
                                             C.byref(o), self._epoch, 5000, vp(out4), st), "orx_xchg_step")
            self._epoch += 3
            return out4[:2]        # already the global (loss, l2_loss): the partials travel through the mailboxes
        ids = torch.cat([uid, pid, nid])
        counts, send_local, slot = eng.owner_bucket_combined(ids, B, self.U, R)
        _lib.check(eng.lib.orx_xchg_push_ids(eng.h, x, vp(counts), vp(send_local), 3 * B, st), "orx_xchg_push_ids")
        self._barrier()                                                # A: every owner has its requests
        _lib.check(eng.lib.orx_xchg_gather_push(eng.h, x, vp(self.table), self.table.shape[0], self.gin_rows,
                                                vp(self._req), vp(self._n_dev), None, st), "orx_xchg_gather_push")
        self._barrier()                                                # B: my rows (and inbox bases) have landed
        out4 = torch.zeros(4, dtype=torch.float32, device=uid.device)
        _lib.check(eng.lib.orx_xchg_grad_push(eng.h, self.kind, x, vp(counts), vp(slot), B, D, self.margin, c_loss, c_l2,
                                              1.0 / (B * R), vp(out4), st), "orx_xchg_grad_push")
        self._barrier()                                                # C: every gradient row is in its owner's inbox
        o = eng.make_opt(self.opt_kind, self.lr, self.eps, self.b1, self.b2, self.iterations)
        tab = eng.make_table(self.table, *self.slots)
        _lib.check(eng.lib.orx_sparse_apply_devn(eng.h, C.byref(tab), vp(self._req), C.c_void_p(self._bufs[3].ptr), W,
                                                 self.gin_rows, vp(self._n_dev), C.byref(o), st),
                   "orx_sparse_apply_devn")
        out = out4[:2].clone()
        if reduce_loss:
            dist.all_reduce(out)
        return out

    def check(self):
        """Raise if a barrier timed out or a gradient inbox overflowed (sticky device flag; one tiny D2H read)."""
        code = int(self._bufs[4].t[self.world].item())
        if code:
            raise RuntimeError({1: "mailbox barrier timed out: a peer rank never arrived",
                                2: "gradient inbox overflow: rebuild with a larger gin_rows"}.get(code, f"error {code}"))

    def close(self):
        torch.cuda.synchronize()
        dist.barrier()
        for p in self._opened:
            self.eng.lib.orx_peer_close(self.eng.h, C.c_void_p(p))
        self._opened = []
        dist.barrier()
        for b in self._bufs:
            b.free()


# ---------------------------------------------------------------------------------------
# bench.py's N>1 leg
# ---------------------------------------------------------------------------------------
def bench(args, rank, world, eng, barrier):
    import bench as B
    from . import native as N
    K, W = args.steps, max(3, args.warmup)
    dev = eng.device
    U, I, D, Bsz = B.U, 12_500_000 * world, B.D, B.B      # BASELINE configs[4]: 100M items x 128 over 8 GPUs
    # Default: NCCL all-to-all exchange.  ORX_SHARDED=peer selects the one-sided NVLink peer-memory step
    # (sharded_peer.py) -- correct, and fast while a rank's shard stays below ~2 GB, but random 512 B rows from a
    # 6.6 GB peer-mapped shard run at 35 GB/s (vs ~600 GB/s at 2 GB; profiles/r1k_p2p_probe.txt), so it is opt-in.
    mode = os.environ.get("ORX_SHARDED", "mailbox")
    use_peer = mode == "peer"
    if mode == "mailbox":
        model = MailboxShardedPairwise(eng, rank, world, U, I, D, Bsz, kind=0, opt_kind=N.ORX_OPT_ADAGRAD, lr=B.LR, seed=1)
    elif use_peer:
        from .sharded_peer import PeerShardedPairwise
        model = PeerShardedPairwise(eng, rank, world, U, I, D, Bsz, kind=0, opt_kind=N.ORX_OPT_ADAGRAD, lr=B.LR, seed=1)
    else:
        model = ShardedPairwise(eng, rank, world, U, I, D, kind=0, opt_kind=N.ORX_OPT_ADAGRAD, lr=B.LR, seed=1)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    host_ids = [tuple(torch.randint(0, n, (Bsz,), generator=g, dtype=torch.int32).pin_memory() for n in (U, I, I))
                for _ in range(B.N_BATCHES)]
    dev_ids = [tuple(x.to(dev) for x in b) for b in host_ids]
    clocks = B.ClockSampler(dev.index or 0) if rank == 0 else None
    for i in range(W):
        model.step(*dev_ids[i % B.N_BATCHES], reduce_loss=False)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for i in range(K):
        model.step(*dev_ids[i % B.N_BATCHES], reduce_loss=False)   # loss stays a per-rank partial on the device
    e1.record()
    barrier()
    t1 = time.time()
    seconds = e0.elapsed_time(e1) * 1e-3
    if clocks:
        clocks.window(t0, t1)
    # e2e: pinned host ids in, global loss out to the host, every step
    last = 0.0
    for i in range(W):
        last = model.step(*(x.to(dev, non_blocking=True) for x in host_ids[i % B.N_BATCHES])).cpu()
    barrier()
    t0 = time.time()
    e0.record()
    pinned = [torch.zeros(2).pin_memory() for _ in range(2)]
    prev = None
    for i in range(K):
        out = model.step(*(x.to(dev, non_blocking=True) for x in host_ids[i % B.N_BATCHES]))   # global loss
        pinned[i & 1].copy_(out, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        if prev is not None:                      # read every step's loss on the host, one step behind
            prev[0].synchronize()
            last = prev[1].clone()
        prev = (ev, pinned[i & 1])
    prev[0].synchronize()
    last = prev[1].clone()
    e1.record()
    barrier()
    t1 = time.time()
    if clocks:
        clocks.window(t0, t1)
    e2e_seconds = e0.elapsed_time(e1) * 1e-3
    if hasattr(model, "check"):
        model.check()
    # NVLink-bound exchange: bytes per GPU per direction per step
    link_bytes = 2.0 * (world - 1) / world * (3 * D + 2) * 4 * Bsz
    nvlink_peak = 770.0
    roofline = {"bound": "nvlink", "achieved": link_bytes / (seconds / K) / 1e9, "peak": nvlink_peak, "unit": "GB/s",
                "frac": link_bytes / (seconds / K) / 1e9 / nvlink_peak, "traffic": None,
                "peak_source": "B200_PROFILING.md measured peer copy 770 GB/s per direction per GPU",
                "algorithmic_bytes_per_launch": link_bytes,
                "note": "per-GPU per-direction NVLink bytes of the row + gradient exchange; the N=1 line carries "
                        "the HBM roofline of the fused kernel"}
    return {"seconds": seconds, "e2e_seconds": e2e_seconds,
            "clocks": clocks.stop() if clocks else None, "launches": model.launches_per_step * K * world,
            "roofline": roofline,
            "e2e_api": f"openrec_b200.{type(model).__module__.split('.')[-1]}.{type(model).__name__}.step"
                       "; pinned host ids in, global loss to host each step",
            "extra": {"last_loss": [float(x) for x in last], "total_items": I, "total_users": U,
                      "exchange": {"peer": "NVLink peer loads/stores inside liborx kernels (CUDA IPC), 2 barriers/step",
                                   "mailbox": "liborx kernels storing ids / rows / gradient rows into the peers' IPC-mapped "
                                              "mailboxes over NVLink, 3 device-side flag barriers, no host sync",
                                   }.get(mode, "NCCL all-to-all (counts, ids, rows, gradient rows)")}}
