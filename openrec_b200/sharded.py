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
import sys
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


class _PeerBuf:
    """A cudaMalloc'd, IPC-exportable device buffer viewed as a torch tensor (orx_peer_alloc)."""

    def __init__(self, eng, n_elems, dtype):
        self.eng = eng
        n = max(int(n_elems), 4)
        self.bytes = n * 4
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        _lib.check(eng.lib.orx_peer_alloc(eng.h, self.bytes, C.byref(ptr), handle), "orx_peer_alloc")
        self.ptr, self.handle = ptr.value, handle.raw
        typestr = {torch.float32: "<f4", torch.int32: "<i4"}[dtype]
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (self.ptr, False),
                                         "version": 2, "strides": None}
        self.t = torch.as_tensor(self, device=eng.device)     # zero-copy view of our own allocation

    def free(self):
        if self.ptr:
            self.t = None
            self.eng.lib.orx_peer_free(self.eng.h, C.c_void_p(self.ptr))
            self.ptr = 0


_MAILBOX_DTYPES = (torch.int32, torch.int32, torch.float32, torch.float32, torch.float32, torch.float32, torch.int32,
                   torch.int32)   # tripbox idbox got gotb gin ginb meta flags
_SHARD_ERRORS = {1: "a peer rank never arrived (flag wait timed out)",
                 2: "more triplets were routed to this rank than home_cap: rebuild with a larger home_cap",
                 3: "one owner received more requests than req_cap", 4: "gradient inbox overflow: rebuild with a larger gin_cap"}


class HomeRoutedPairwise:
    """Row-sharded BPR (kind 0) / UCML (kind 1) step, "home-routed" (csrc/orx_shard.cu): row r of the user and item
    tables lives on rank ``r % R``; a triplet is computed on the rank that owns its USER row, so only the two item rows
    travel in and the two item gradient rows travel out, as peer stores of aligned rows into IPC-mapped mailboxes.  One
    C call per step (six launches, flag words in peer memory instead of barriers, no collective, no host sync).
    ``torch.distributed`` is used once, to swap the 64-byte IPC handles.

    ``peers=None``: the R ranks are R processes (one per GPU) and the mailboxes are exchanged over ``torch.distributed``.
    ``peers=<LoopbackGroup>``: R virtual ranks share one device and one stream (1-GPU parity test of the same kernels).
    """

    def __init__(self, eng, rank, world, total_users, total_items, dim, batch, *, kind=0, opt_kind=1, lr=0.05, eps=1e-7,
                 beta1=0.9, beta2=0.999, margin=0.5, seed=0, init=True, home_cap=None, gin_cap=None, timeout_ms=20000,
                 peers=None, tables=None, slots=None):
        if dim % 4 or dim > 512:
            raise ValueError("the sharded step needs dim % 4 == 0 and dim <= 512")
        if opt_kind not in (0, 1, 2):
            raise ValueError("the sharded step supports SGD, Adagrad and row-sparse Adam")
        self.eng, self.rank, self.world = eng, rank, world
        self.U, self.I, self.D, self.B = total_users, total_items, dim, batch
        self.kind, self.opt_kind, self.lr, self.eps, self.b1, self.b2, self.margin = kind, opt_kind, lr, eps, beta1, beta2, margin
        self.iterations = 0
        dev = eng.device
        self.ru = (total_users - rank + world - 1) // world     # rows r with r % world == rank
        self.ri = (total_items - rank + world - 1) // world
        if tables is not None:          # shards owned by the caller (openrec.tf2.recommenders.ShardedBPR: keras variables)
            self.user, self.item, self.bias = tables
            want = ((max(self.ru, 1), dim), (max(self.ri, 1), dim), (max(self.ri, 1), 1))
            if tuple(tuple(t.shape) for t in tables) != want:
                raise ValueError(f"shard shapes {[tuple(t.shape) for t in tables]} != {want}")
        else:
            self.user = torch.zeros(max(self.ru, 1), dim, dtype=torch.float32, device=dev)
            self.item = torch.zeros(max(self.ri, 1), dim, dtype=torch.float32, device=dev)
            self.bias = torch.zeros(max(self.ri, 1), 1, dtype=torch.float32, device=dev)
            if init:
                for k, t in enumerate((self.user, self.item, self.bias)):
                    eng.fill_uniform(t, -0.05, 0.05, seed * 1000003 + rank * 17 + k)
        n_slots = {0: 0, 1: 1, 2: 2}[opt_kind]
        fill = 0.1 if opt_kind == 1 else 0.0
        mk = lambda t: [torch.full_like(t, fill) for _ in range(n_slots)] + [None] * (2 - n_slots)
        if slots is not None:           # optimizer slots owned by the caller (the keras optimizer's slot tensors)
            self.user_slots, self.item_slots, self.bias_slots = (list(x) for x in slots)
        else:
            self.user_slots, self.item_slots, self.bias_slots = mk(self.user), mk(self.item), mk(self.bias)
        # capacities: worst case for small problems (tests), 2x the expected load for big ones (errors are sticky flags)
        small = world * batch <= (1 << 16)
        self.home_cap = int(home_cap or (world * batch if small else 2 * batch))
        self.req_cap = 2 * self.home_cap
        self.gin_cap = int(gin_cap or (world * (2 * self.home_cap + 32) if small else 4 * batch + 32 * world))
        self.gin_cap = (self.gin_cap + 31) // 32 * 32
        self._x = _lib.OrxShard(world, rank, dim, batch, self.home_cap, self.req_cap, self.gin_cap, timeout_ms,
                                0, 0, 0, 0, 0, 0, 0, 0)
        sizes = (C.c_int64 * 8)()
        _lib.check(eng.lib.orx_shard_sizes(C.byref(self._x), sizes), "orx_shard_sizes")
        self._loop = peers
        self._opened = []
        if peers is None:
            self._bufs = [_PeerBuf(eng, sizes[k], _MAILBOX_DTYPES[k]) for k in range(8)]
            mine = [b.ptr for b in self._bufs]
            torch.cuda.synchronize()
            everyone = [None] * world
            dist.all_gather_object(everyone, [b.handle for b in self._bufs])
            ptrs = np.zeros((8, world), dtype=np.int64)
            for r in range(world):
                for k in range(8):
                    if r == rank:
                        ptrs[k, r] = mine[k]
                    else:
                        p = C.c_void_p()
                        _lib.check(eng.lib.orx_peer_open(eng.h, everyone[r][k], C.byref(p)), "orx_peer_open")
                        self._opened.append(p.value)
                        ptrs[k, r] = p.value
            self._set_pointers(ptrs)
            dist.barrier()
        else:                    # loopback: plain device memory, the group wires the pointer tables
            self._bufs = None
            self._tensors = [torch.zeros(max(int(sizes[k]), 4), dtype=_MAILBOX_DTYPES[k], device=dev) for k in range(8)]
            peers._register(self)
        self._out = torch.zeros(16, 4, dtype=torch.float32, device=dev)     # ring of step outputs
        self._tabs = (eng.make_table(self.user, *self.user_slots), eng.make_table(self.item, *self.item_slots),
                      eng.make_table(self.bias, *self.bias_slots))
        self.launches_per_step = 6      # 4 when every step announces the next batch (step(next_ids=...))
        self._announced = None

    def _set_pointers(self, ptrs):
        self._ptrs = torch.from_numpy(np.ascontiguousarray(ptrs)).to(self.eng.device)
        base = self._ptrs.data_ptr()
        for k, name in enumerate(("tripbox", "idbox", "got", "gotb", "gin", "ginb", "meta", "flags")):
            setattr(self._x, name, base + 8 * self.world * k)

    def _flags(self):
        return self._bufs[7].t if self._bufs is not None else self._tensors[7]

    def _call(self, uid, pid, nid, c_loss, c_l2, lo, hi, nxt=None, epoch=None):
        eng = self.eng
        B = uid.numel()
        if B > self.B:
            raise ValueError("batch larger than the mailboxes this model was built for")
        epoch = self.iterations if epoch is None else epoch
        out4 = self._out[epoch % 16]
        o = eng.make_opt(self.opt_kind, self.lr, self.eps, self.b1, self.b2, epoch)
        vp = lambda t: C.c_void_p(t.data_ptr())
        nx = (vp(nxt[0]), vp(nxt[1]), vp(nxt[2]), nxt[0].numel()) if nxt is not None else (None, None, None, 0)
        _lib.check(eng.lib.orx_shard_step(eng.h, self.kind, C.byref(self._x), C.byref(self._tabs[0]), C.byref(self._tabs[1]),
                                          C.byref(self._tabs[2]), vp(uid), vp(pid), vp(nid), B, *nx, self.U, self.I, self.margin,
                                          c_loss, c_l2, 1.0 / (B * self.world), C.byref(o), epoch, lo, hi, vp(out4),
                                          eng.stream()), "orx_shard_step")
        return out4

    def step(self, uid, pid, nid, c_loss=1.0, c_l2=1.0, reduce_loss=True, next_ids=None):
        """uid/pid/nid: this rank's int32 GLOBAL ids on the device (every rank must pass the same batch size).
        Returns a [2] device tensor = the GLOBAL (loss, l2_loss): the partials ride the meta mailboxes.

        ``next_ids=(uid, pid, nid)`` announces the batch of the NEXT step (device tensors; the next call must pass these
        very tensors): its routing and request phases -- two of the four cross-rank handoffs of a step -- then run inside
        this step's apply launch, and the next step is four launches instead of six.  All ranks announce, or none."""
        if self._loop is not None:
            raise RuntimeError("loopback ranks are stepped by their LoopbackGroup")
        if self._announced is not None and any(a.data_ptr() != b.data_ptr() for a, b in zip(self._announced, (uid, pid, nid))):
            raise ValueError("this batch is not the one announced with next_ids= in the previous step")
        self.iterations += 1
        out = self._call(uid, pid, nid, c_loss, c_l2, 0, 5, nxt=next_ids)[:2]
        self._announced = tuple(next_ids) if next_ids is not None else None     # also keeps the tensors alive
        return out

    def check(self):
        """Raise if a flag wait timed out or a mailbox overflowed (sticky device word; one tiny D2H read)."""
        code = int(self._flags()[4 * 64].item())
        if code:
            raise RuntimeError("sharded step: " + _SHARD_ERRORS.get(code, f"error {code}"))

    # ---- global <-> shard (tests, checkpoints)
    def load_global(self, user, item, bias):
        r, R = self.rank, self.world
        dev = self.eng.device
        self.user[:self.ru] = torch.as_tensor(np.ascontiguousarray(user[r::R]), dtype=torch.float32).to(dev)
        self.item[:self.ri] = torch.as_tensor(np.ascontiguousarray(item[r::R]), dtype=torch.float32).to(dev)
        self.bias[:self.ri] = torch.as_tensor(np.ascontiguousarray(bias[r::R]), dtype=torch.float32).reshape(-1, 1).to(dev)

    def local_shards(self):
        return self.user[:self.ru], self.item[:self.ri], self.bias[:self.ri]

    def gather_global(self):
        """-> (user, item, bias) full tables on every rank (test helper; sizes must be small)."""
        outs = []
        for t, total in zip(self.local_shards(), (self.U, self.I, self.I)):
            per = (total + self.world - 1) // self.world
            pad = torch.zeros(per, t.shape[1], dtype=t.dtype, device=t.device)
            pad[:t.shape[0]] = t
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad)
            outs.append(torch.stack(parts, 1).reshape(per * self.world, t.shape[1])[:total])   # row = local*R + rank
        return outs

    # ---- per-rank shard checkpoint (SURVEY 8f N4 for tables that only exist sharded)
    def save_shard(self, path):
        """Write this rank's rows + optimizer slots + step count to ``path`` (one .npz per rank)."""
        arrs = {"meta": np.array([self.rank, self.world, self.U, self.I, self.D, self.kind, self.opt_kind, self.iterations],
                                 dtype=np.int64)}
        for name, t, slots in (("user", self.user[:self.ru], self.user_slots), ("item", self.item[:self.ri], self.item_slots),
                               ("bias", self.bias[:self.ri], self.bias_slots)):
            arrs[name] = t.cpu().numpy()
            for k, s in enumerate(slots):
                if s is not None:
                    arrs[f"{name}_s{k}"] = s[:t.shape[0]].cpu().numpy()
        np.savez(path, **arrs)

    def load_shard(self, path):
        z = np.load(path)
        meta = z["meta"].tolist()
        want = [self.rank, self.world, self.U, self.I, self.D, self.kind, self.opt_kind]
        if meta[:7] != want:
            raise ValueError(f"shard checkpoint {path} was written for (rank, world, U, I, D, kind, opt) = {meta[:7]}, "
                             f"this model is {want}")
        dev = self.eng.device
        for name, t, slots in (("user", self.user[:self.ru], self.user_slots), ("item", self.item[:self.ri], self.item_slots),
                               ("bias", self.bias[:self.ri], self.bias_slots)):
            t.copy_(torch.from_numpy(z[name]).to(dev))
            for k, s in enumerate(slots):
                if s is not None:
                    s[:t.shape[0]].copy_(torch.from_numpy(z[f"{name}_s{k}"]).to(dev))
        self.iterations = int(meta[7])

    def close(self):
        torch.cuda.synchronize()
        if self._bufs is None:
            return
        dist.barrier()
        for p in self._opened:
            self.eng.lib.orx_peer_close(self.eng.h, C.c_void_p(p))
        self._opened = []
        dist.barrier()
        for b in self._bufs:
            b.free()
        self._bufs = None


class LoopbackGroup:
    """R virtual ranks of HomeRoutedPairwise on ONE device and ONE stream: every rank has its own liborx handle, tables
    and mailboxes (plain device memory); a step issues phase k of the six launches for every rank before phase k + 1,
    so every flag a kernel waits for is already set.  Same kernels, same peer-pointer tables as the multi-GPU step."""

    def __init__(self, world, total_users, total_items, dim, batch, device=None, **kw):
        from . import native
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.world = world
        self.ranks = []
        self._announced = None
        self._engines = [native.Engine(dev.index or 0) for _ in range(world)]
        for r in range(world):
            HomeRoutedPairwise(self._engines[r], r, world, total_users, total_items, dim, batch, peers=self, **kw)
        ptrs = np.zeros((8, world), dtype=np.int64)
        for r, m in enumerate(self.ranks):
            for k in range(8):
                ptrs[k, r] = m._tensors[k].data_ptr()
        for m in self.ranks:
            m._set_pointers(ptrs)
        torch.cuda.synchronize()

    def _register(self, m):
        self.ranks.append(m)

    def step(self, batches, c_loss=1.0, c_l2=1.0, next_batches=None):
        """batches[r] = (uid, pid, nid) of rank r.  Returns the [2] global (loss, l2_loss) tensor of every rank.
        ``next_batches`` announces the next step's batches: their route / request phases are issued between compute and
        apply of this step -- where the multi-GPU step runs them (inside the apply launch; fused roles of R virtual ranks
        on one stream would wait for each other, so the loopback issues the stand-alone kernels at that point)."""
        if self._announced is not None:
            for b, a in zip(batches, self._announced):
                if any(x.data_ptr() != y.data_ptr() for x, y in zip(b, a)):
                    raise ValueError("these batches are not the ones announced in the previous step")
        for m in self.ranks:
            m.iterations += 1
        outs = [None] * self.world
        for ph in range(6):       # phases 0 / 1 of an announced batch were issued a step ago: the C call skips them
            fused = ph == 4 and next_batches is not None and self.world == 1    # one rank: the real fused launch
            if ph == 4 and next_batches is not None and not fused:
                for early in (0, 1):
                    for r, m in enumerate(self.ranks):
                        m._call(*next_batches[r], c_loss, c_l2, early, early, epoch=m.iterations + 1)
            for r, m in enumerate(self.ranks):
                outs[r] = m._call(*batches[r], c_loss, c_l2, ph, ph, nxt=next_batches[r] if fused else None)
        self._announced = [tuple(b) for b in next_batches] if next_batches is not None else None
        return [o[:2] for o in outs]

    def load_global(self, user, item, bias):
        for m in self.ranks:
            m.load_global(user, item, bias)

    def gather_global(self):
        R = self.world
        outs = []
        for k, total in enumerate((self.ranks[0].U, self.ranks[0].I, self.ranks[0].I)):
            shards = [m.local_shards()[k] for m in self.ranks]
            full = torch.zeros(total, shards[0].shape[1], dtype=torch.float32, device=shards[0].device)
            for r in range(R):
                full[r::R] = shards[r]
            outs.append(full)
        return outs

    def check(self):
        for m in self.ranks:
            m.check()

    def close(self):
        torch.cuda.synchronize()
        for e in self._engines:
            e.close()


# ---------------------------------------------------------------------------------------
# bench.py's N>1 leg
# ---------------------------------------------------------------------------------------
def bench(args, rank, world, eng, barrier, clocks=None):
    import bench as B
    from . import native as N
    K, W = args.steps, max(3, args.warmup)
    dev = eng.device
    U, I, D, Bsz = B.U, 12_500_000 * world, B.D, B.B      # BASELINE configs[4]: 100M items x 128 over 8 GPUs
    # Default: the home-routed peer-store exchange (csrc/orx_shard.cu); ORX_SHARDED=nccl selects the NCCL all-to-all form
    mode = os.environ.get("ORX_SHARDED", "home")
    if mode == "home":
        model = HomeRoutedPairwise(eng, rank, world, U, I, D, Bsz, kind=0, opt_kind=N.ORX_OPT_ADAGRAD, lr=B.LR, seed=1)
    else:
        model = ShardedPairwise(eng, rank, world, U, I, D, kind=0, opt_kind=N.ORX_OPT_ADAGRAD, lr=B.LR, seed=1)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    host_ids = [tuple(torch.randint(0, n, (Bsz,), generator=g, dtype=torch.int32).pin_memory() for n in (U, I, I))
                for _ in range(B.N_BATCHES)]
    dev_ids = [tuple(x.to(dev) for x in b) for b in host_ids]

    # Every step announces the next batch (HomeRoutedPairwise.step(next_ids=...)): a training loop knows it -- the data
    # pipeline is a step ahead -- and the routing + request phases of step t+1 then hide under the apply phase of step t.
    # ORX_SHARD_ANNOUNCE=0 measures the six-launch step without it.
    announce = mode == "home" and os.environ.get("ORX_SHARD_ANNOUNCE", "1") != "0"
    NB = B.N_BATCHES
    cnt = {"k": 0}

    def step(i):
        k = cnt["k"]
        cnt["k"] = k + 1
        if announce:
            model.step(*dev_ids[k % NB], reduce_loss=False, next_ids=dev_ids[(k + 1) % NB])
        else:
            model.step(*dev_ids[k % NB], reduce_loss=False)

    def drain():            # the batch announced by the last step of a loop is stepped outside the timed region
        if announce and model._announced is not None:
            model.step(*model._announced)
            cnt["k"] += 1

    for i in range(W):
        step(i)
    seconds = B._timed(step, K, barrier, torch, clocks)
    drain()
    # e2e: pinned host ids in, global loss out to the host, every step (read one step behind, so the copy overlaps).
    # The ids of batch k+2 are uploaded on a side stream during step k into a 4-deep ring of device buffers.
    pinned = [torch.zeros(2).pin_memory() for _ in range(2)]
    state = {"prev": None, "last": None}
    ring = [tuple(torch.empty(Bsz, dtype=torch.int32, device=dev) for _ in range(3)) for _ in range(4)]
    up_ev, done_ev = [None] * 4, [None] * 4
    side = torch.cuda.Stream(device=dev)
    e2e = {"k": 0}

    def upload(k):
        if done_ev[k % 4] is not None:
            side.wait_event(done_ev[k % 4])          # the step that last read this ring slot (batch k - 4) has finished
        with torch.cuda.stream(side):
            for dst, src in zip(ring[k % 4], host_ids[k % NB]):
                dst.copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(side)
        up_ev[k % 4] = ev

    def e2e_step(i):
        k = e2e["k"]
        e2e["k"] = k + 1
        if k == 0:
            upload(0)
            upload(1)
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(up_ev[k % 4])
        cur.wait_event(up_ev[(k + 1) % 4])
        out = model.step(*ring[k % 4], next_ids=ring[(k + 1) % 4]) if announce else model.step(*ring[k % 4])   # global (loss, l2)
        d = torch.cuda.Event()
        d.record()
        done_ev[k % 4] = d
        upload(k + 2)
        pinned[i & 1].copy_(out, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        if state["prev"] is not None:
            state["prev"][0].synchronize()
            state["last"] = state["prev"][1].clone()
        state["prev"] = (ev, pinned[i & 1])

    for i in range(W):
        e2e_step(i)
    e2e_seconds = B._timed(e2e_step, K, barrier, torch, clocks)
    drain()
    state["prev"][0].synchronize()
    last = state["prev"][1].clone()
    if hasattr(model, "check"):
        model.check()
    checked = None
    if getattr(args, "check", False):
        # parity at the full table shape (BASELINE configs[4]); the oracle stays under tests/ (test infrastructure)
        tests_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")
        if tests_dir not in sys.path:
            sys.path.insert(0, tests_dir)
        import shard_check
        access = shard_check.HomeRoutedAccess(model) if mode == "home" else shard_check.CombinedAccess(model)
        checked = shard_check.run(model, access, rank, world, U, I, D, Bsz, kind=0, opt_kind=1, lr=B.LR)
    # NVLink-bound exchange: bytes per GPU per direction per step: rows out as owner + gradient rows out as home
    rows_each_way = 2 if mode == "home" else 3
    link_bytes = 2.0 * (world - 1) / world * rows_each_way * (D + 1) * 4 * Bsz
    nvlink_peak = 770.0
    ms = seconds / K * 1e3
    roofline = {"bound": "nvlink", "kernel": "k_sh_serve + k_sh_compute (peer stores)" if mode == "home" else "NCCL all-to-all",
                "achieved": link_bytes / (seconds / K) / 1e9, "peak": nvlink_peak, "unit": "GB/s",
                "frac": link_bytes / (seconds / K) / 1e9 / nvlink_peak, "traffic": None,
                "peak_source": "B200_PROFILING.md measured peer copy 770 GB/s per direction per GPU",
                "algorithmic_bytes_per_launch": link_bytes,
                "note": f"per-GPU per-direction NVLink bytes of the item-row + gradient-row exchange ({rows_each_way} rows each way "
                        f"per triplet, (N-1)/N of them remote) over the WHOLE step time ({ms:.3f} ms): the step also does the local "
                        "HBM work of the single-GPU step; the N = 1 line carries the HBM roofline of the fused kernel"}
    per_step = 4 if announce else model.launches_per_step
    return {"seconds": seconds, "e2e_seconds": e2e_seconds, "launches": per_step * K * world,
            "units_per_step": Bsz, "h2d": 3 * 4 * Bsz, "d2h": 8, "roofline": roofline,
            "e2e_api": f"openrec_b200.sharded.{type(model).__name__}.step; pinned host ids in, global loss to host each step",
            "extra": {"last_loss": [float(x) for x in last], "total_items": I, "total_users": U, "check": checked,
                      "exchange": {"home": "home-routed: triplets computed on the user row's owner; item rows / gradient rows "
                                           "as peer stores into IPC-mapped mailboxes over NVLink, flag words instead of "
                                           "barriers, no collective, no host sync",
                                   }.get(mode, "NCCL all-to-all (counts, ids, rows, gradient rows)")}}
