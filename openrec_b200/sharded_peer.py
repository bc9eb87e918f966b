"""One-sided row-sharded BPR / UCML step over NVLink peer memory (SURVEY 8e "B200-native fused design").

Same partitioning as ``sharded.py`` (row r on rank ``r % R``, combined local table ``[user rows | item rows]``),
but no NCCL in the data path: every rank maps every other rank's shard and inbox through CUDA IPC, and one liborx
kernel per rank gathers rows with direct peer loads and pushes gradient rows with direct peer stores
(``orx_peer_pairwise_push``); owners then deduplicate and apply (``orx_peer_apply``).  ``torch.distributed`` is
used only to exchange the 64-byte IPC handles once and for the two tiny barriers of a step.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import OrxPeer, OrxTable


class _PeerBuf:
    """A cudaMalloc'd, IPC-exportable device buffer viewed as a torch tensor."""

    def __init__(self, eng, shape, dtype):
        self.eng, self.shape, self.dtype = eng, tuple(shape), dtype
        n = int(np.prod(self.shape))
        self.bytes = max(n * torch.empty((), dtype=dtype).element_size(), 16)
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        _lib.check(eng.lib.orx_peer_alloc(eng.h, self.bytes, C.byref(ptr), handle), "orx_peer_alloc")
        self.ptr, self.handle = ptr.value, handle.raw
        typestr = {torch.float32: "<f4", torch.int32: "<i4"}[dtype]
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": typestr, "data": (self.ptr, False),
                                         "version": 2, "strides": None}
        self.t = torch.as_tensor(self, device=eng.device)     # zero-copy view of our own allocation

    def free(self):
        if self.ptr:
            self.t = None
            self.eng.lib.orx_peer_free(self.eng.h, C.c_void_p(self.ptr))
            self.ptr = 0


class PeerShardedPairwise:
    """BPR (kind 0) / UCML (kind 1) over peer memory.  dim must be 128 or 256; optimizer SGD / Adagrad / lazy Adam."""

    def __init__(self, eng, rank, world, total_users, total_items, dim, batch, *, kind=0, opt_kind=1, lr=0.05, eps=1e-7,
                 beta1=0.9, beta2=0.999, margin=0.5, seed=0, init=True):
        self.eng, self.rank, self.world = eng, rank, world
        self.U, self.I, self.D, self.B = total_users, total_items, dim, batch
        self.kind, self.opt_kind, self.lr, self.eps, self.b1, self.b2, self.margin = kind, opt_kind, lr, eps, beta1, beta2, margin
        self.iterations = 0
        dev = eng.device
        self.ru = (total_users - rank + world - 1) // world
        self.ri = (total_items - rank + world - 1) // world
        rows = self.ru + self.ri
        self.cap = 3 * batch
        self._emb = _PeerBuf(eng, (rows, dim), torch.float32)
        self._bias = _PeerBuf(eng, (rows,), torch.float32)
        self._in_emb = _PeerBuf(eng, (world * self.cap, dim), torch.float32)
        self._in_bias = _PeerBuf(eng, (world * self.cap,), torch.float32)
        self._in_ids = _PeerBuf(eng, (world * self.cap,), torch.int32)
        self._in_cnt = _PeerBuf(eng, (world,), torch.int32)
        self.emb, self.bias = self._emb.t, self._bias.t
        if init:
            eng.fill_uniform(self.emb, -0.05, 0.05, seed * 1000003 + rank * 17)
            tmp = torch.empty(rows, dtype=torch.float32, device=dev)
            eng.fill_uniform(tmp, -0.05, 0.05, seed * 1000003 + rank * 17 + 7)
            tmp[:self.ru] = 0.0
            self.bias.copy_(tmp)
        n_slots = {0: 0, 1: 1, 2: 2}[opt_kind]
        fill = 0.1 if opt_kind == 1 else 0.0
        self.emb_slots = [torch.full_like(self.emb, fill) for _ in range(n_slots)] + [None] * (2 - n_slots)
        self.bias_slots = [torch.full_like(self.bias, fill) for _ in range(n_slots)] + [None] * (2 - n_slots)
        # exchange IPC handles, map the peers
        mine = [b.handle for b in (self._emb, self._bias, self._in_emb, self._in_bias, self._in_ids, self._in_cnt)]
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        self._opened = []
        ptrs = np.zeros((6, world), dtype=np.int64)
        for r in range(world):
            for k in range(6):
                if r == rank:
                    ptrs[k, r] = (self._emb, self._bias, self._in_emb, self._in_bias, self._in_ids, self._in_cnt)[k].ptr
                else:
                    p = C.c_void_p()
                    _lib.check(eng.lib.orx_peer_open(eng.h, everyone[r][k], C.byref(p)), "orx_peer_open")
                    self._opened.append(p.value)
                    ptrs[k, r] = p.value
        self._ptrs = torch.from_numpy(ptrs).to(dev)            # [6, world] device pointer table
        base = self._ptrs.data_ptr()
        self._peer = OrxPeer(world, rank, dim, 0, total_users, total_items, self.cap,
                             *[base + 8 * world * k for k in range(6)])
        self._pos = torch.empty(3 * batch, dtype=torch.int32, device=dev)
        self._flag = torch.zeros(1, dtype=torch.float32, device=dev)
        self.launches_per_step = 3 + 2 + 3      # hist/publish/positions, peer step + reduce, inbox index/apply/tail
        dist.barrier()

    def _barrier(self):
        dist.all_reduce(self._flag)             # stream-ordered: every rank's previous work is complete past this point

    def _tab(self, var, slots):
        s0, s1 = slots
        return OrxTable(var.data_ptr(), s0.data_ptr() if s0 is not None else None,
                        s1.data_ptr() if s1 is not None else None, var.shape[0], var.shape[1] if var.dim() == 2 else 1)

    def step(self, uid, pid, nid, c_loss=1.0, c_l2=1.0, reduce_loss=True):
        eng, R = self.eng, self.world
        B = uid.numel()
        if B > self.B:
            raise ValueError("batch larger than the inbox capacity this model was built for")
        self.iterations += 1
        out4 = torch.zeros(4, dtype=torch.float32, device=uid.device)
        self._barrier()                                            # shards final, inboxes consumed
        _lib.check(eng.lib.orx_peer_pairwise_push(eng.h, self.kind, C.byref(self._peer), C.c_void_p(uid.data_ptr()),
                                                  C.c_void_p(pid.data_ptr()), C.c_void_p(nid.data_ptr()), B,
                                                  C.c_void_p(self._pos.data_ptr()), self.margin, c_loss, c_l2,
                                                  1.0 / (B * R), C.c_void_p(out4.data_ptr()), eng.stream()),
                   "orx_peer_pairwise_push")
        self._barrier()                                            # every rank's pushes have landed
        o = eng.make_opt(self.opt_kind, self.lr, self.eps, self.b1, self.b2, self.iterations)
        _lib.check(eng.lib.orx_peer_apply(eng.h, C.byref(self._tab(self.emb, self.emb_slots)),
                                          C.byref(self._tab(self.bias.reshape(-1, 1), self.bias_slots)),
                                          C.c_void_p(self._in_ids.ptr), C.c_void_p(self._in_emb.ptr),
                                          C.c_void_p(self._in_bias.ptr), C.c_void_p(self._in_cnt.ptr), R, self.cap,
                                          C.byref(o), eng.stream()), "orx_peer_apply")
        out = out4[:2].clone()
        if reduce_loss:
            dist.all_reduce(out)
        return out

    # ---- test helpers (same contract as sharded.ShardedPairwise)
    def load_global(self, user, item, bias):
        r, R = self.rank, self.world
        self.emb[:self.ru] = torch.as_tensor(user[r::R], dtype=torch.float32)
        self.emb[self.ru:] = torch.as_tensor(item[r::R], dtype=torch.float32)
        self.bias[:self.ru] = 0.0
        self.bias[self.ru:] = torch.as_tensor(bias[r::R], dtype=torch.float32).reshape(-1)

    def gather_global(self):
        outs = []
        for t, total in ((self.emb[:self.ru], self.U), (self.emb[self.ru:], self.I),
                         (self.bias[self.ru:].reshape(-1, 1), self.I)):
            per = (total + self.world - 1) // self.world
            pad = torch.zeros(per, t.shape[1], dtype=t.dtype, device=t.device)
            pad[:t.shape[0]] = t
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad)
            outs.append(torch.stack(parts, 1).reshape(per * self.world, t.shape[1])[:total])
        return outs

    def close(self):
        torch.cuda.synchronize()
        dist.barrier()
        for p in self._opened:
            self.eng.lib.orx_peer_close(self.eng.h, C.c_void_p(p))
        self._opened = []
        dist.barrier()
        for b in (self._emb, self._bias, self._in_emb, self._in_bias, self._in_ids, self._in_cnt):
            b.free()
