"""Row-sharded BPR / UCML step across the GPUs of one NVSwitch box (BASELINE configs[4]).

The reference is single-device; this is the scale-out of the same synchronous step:
row r of every table lives on rank ``r % R`` at local row ``r // R`` (optimizer slots alongside);
each rank owns B triplets of the global batch.  One step =

  1. orx_owner_bucket      : sort this rank's 3B lookups by owner                      (liborx)
  2. all-to-all            : lookup counts, then local-row ids, to the owners          (NCCL over NVLink)
  3. orx_gather            : owners read the requested rows from their shard          (liborx)
  4. all-to-all            : rows back to the requesters
  5. orx_pairwise_grad_slots: score, loss and per-lookup gradient rows, in place of the
                             rows' slots (pre-step values everywhere: nothing has been written yet)
  6. all-to-all            : gradient rows to the owners
  7. orx_sparse_apply      : owners sum duplicates (across ALL ranks' lookups) and apply the
                             optimizer once per unique row                             (liborx)
  8. all-reduce            : (loss, l2_loss)

``torch.distributed`` is plumbing (one process per GPU, NCCL; gloo in the CPU logic tests); every
arithmetic op is a liborx kernel reached through ``eng``.
"""
from __future__ import annotations

import json
import os
import time

import torch
import torch.distributed as dist

from . import _lib


def _a2a(out, inp, out_splits, in_splits):
    dist.all_to_all_single(out, inp, output_split_sizes=out_splits, input_split_sizes=in_splits)


class ShardedPairwise:
    """BPR (kind 0) / UCML (kind 1) with row-sharded tables.  ``eng`` is a native.Engine (or the
    oracle-backed stand-in of tests/fake_engine.py on CPU)."""

    def __init__(self, eng, rank, world, total_users, total_items, dim, *, kind=0, opt_kind=1, lr=0.05, eps=1e-7,
                 beta1=0.9, beta2=0.999, margin=0.5, seed=0, init=True):
        self.eng, self.rank, self.world = eng, rank, world
        self.U, self.I, self.D = total_users, total_items, dim
        self.kind, self.opt_kind, self.lr, self.eps, self.b1, self.b2, self.margin = kind, opt_kind, lr, eps, beta1, beta2, margin
        self.iterations = 0
        dev = eng.device
        ru = (total_users - rank + world - 1) // world     # rows r with r % world == rank
        ri = (total_items - rank + world - 1) // world
        self.user = torch.empty(ru, dim, dtype=torch.float32, device=dev)
        self.item = torch.empty(ri, dim, dtype=torch.float32, device=dev)
        self.bias = torch.empty(ri, 1, dtype=torch.float32, device=dev)
        if init:
            for k, t in enumerate((self.user, self.item, self.bias)):
                eng.fill_uniform(t, -0.05, 0.05, seed * 1000003 + rank * 17 + k)
        n_slots = {0: 0, 1: 1, 2: 2, 3: 2}[opt_kind]
        fill = 0.1 if opt_kind == 1 else 0.0
        self.slots = [[torch.full_like(t, fill) for _ in range(n_slots)] + [None] * (2 - n_slots)
                      for t in (self.user, self.item, self.bias)]
        self.launches_per_step = 3 * 3 + 3 + 2 + 3 * 2   # buckets, gathers, grad+reduce, sparse applies (+tails)

    def _tables(self):
        return [self.eng.make_table(t, *s) for t, s in zip((self.user, self.item, self.bias), self.slots)]

    def step(self, uid, pid, nid, c_loss=1.0, c_l2=1.0):
        """uid/pid/nid: this rank's int32 GLOBAL ids on the device.  Returns a [2] tensor
        (global loss, global l2_loss) on the device."""
        eng, R, D = self.eng, self.world, self.D
        B = uid.numel()
        dev = uid.device
        self.iterations += 1
        items = torch.cat([pid, nid])
        cu, send_u, slot_u = eng.owner_bucket(uid, R)
        ci, send_i, slot_i = eng.owner_bucket(items, R)
        counts = torch.stack([cu, ci], 1).contiguous()
        rcounts = torch.empty_like(counts)
        dist.all_to_all_single(rcounts, counts)
        host = torch.cat([counts, rcounts]).cpu()                       # the step's one host sync
        su, si = host[:R, 0].tolist(), host[:R, 1].tolist()
        ru, ri = host[R:, 0].tolist(), host[R:, 1].tolist()
        req_u = torch.empty(sum(ru), dtype=torch.int32, device=dev)
        req_i = torch.empty(sum(ri), dtype=torch.int32, device=dev)
        _a2a(req_u, send_u, ru, su)
        _a2a(req_i, send_i, ri, si)
        # owners: fetch rows of their shard, send them back
        rows_u, rows_i, rows_b = eng.gather(self.user, req_u), eng.gather(self.item, req_i), eng.gather(self.bias, req_i)
        got_u = torch.empty(B, D, dtype=torch.float32, device=dev)
        got_i = torch.empty(2 * B, D, dtype=torch.float32, device=dev)
        got_b = torch.empty(2 * B, 1, dtype=torch.float32, device=dev)
        _a2a(got_u, rows_u, su, ru)
        _a2a(got_i, rows_i, si, ri)
        _a2a(got_b, rows_b, si, ri)
        # requesters: gradients of this rank's triplets, written over the rows' slots
        out4 = torch.zeros(4, dtype=torch.float32, device=dev)
        d_u, d_i, d_b = torch.empty_like(got_u), torch.empty_like(got_i), torch.empty_like(got_b)
        eng.pairwise_grad_slots(self.kind, got_u, got_i, got_b, slot_u, slot_i[:B].contiguous(),
                                slot_i[B:].contiguous(), 1.0 / (B * R), d_u, d_i, d_b, out4, self.margin, c_loss, c_l2)
        # gradients to the owners
        g_u, g_i, g_b = torch.empty_like(rows_u), torch.empty_like(rows_i), torch.empty_like(rows_b)
        _a2a(g_u, d_u, ru, su)
        _a2a(g_i, d_i, ri, si)
        _a2a(g_b, d_b, ri, si)
        # owners: dedup across every rank's lookups, optimizer once per unique row
        o = eng.make_opt(self.opt_kind, self.lr, self.eps, self.b1, self.b2, self.iterations)
        tu, ti, tb = self._tables()
        eng.sparse_apply(tu, req_u, g_u, o)
        eng.sparse_apply(ti, req_i, g_i, o)
        eng.sparse_apply(tb, req_i, g_b, o)
        out = out4[:2].clone()
        dist.all_reduce(out)
        return out

    # ---- helpers for tests: assemble / scatter the global tables
    def load_global(self, user, item, bias):
        r, R = self.rank, self.world
        self.user.copy_(torch.as_tensor(user[r::R], dtype=torch.float32))
        self.item.copy_(torch.as_tensor(item[r::R], dtype=torch.float32))
        self.bias.copy_(torch.as_tensor(bias[r::R], dtype=torch.float32))

    def gather_global(self):
        """-> (user, item, bias) full tables on every rank (test helper; sizes must be small)."""
        outs = []
        for t, total in ((self.user, self.U), (self.item, self.I), (self.bias, self.I)):
            per = (total + self.world - 1) // self.world
            pad = torch.zeros(per, t.shape[1], dtype=t.dtype, device=t.device)
            pad[:t.shape[0]] = t
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad)
            full = torch.stack(parts, 1).reshape(per * self.world, t.shape[1])[:total]   # row = local*R + rank
            outs.append(full)
        return outs


# ---------------------------------------------------------------------------------------
# bench.py's N>1 leg
# ---------------------------------------------------------------------------------------
def bench(args, rank, world, eng, barrier):
    import bench as B
    from . import native as N
    K, W = args.steps, max(3, args.warmup)
    dev = eng.device
    U, I, D, Bsz = B.U, 12_500_000 * world, B.D, B.B      # BASELINE configs[4]: 100M items x 128 over 8 GPUs
    model = ShardedPairwise(eng, rank, world, U, I, D, kind=0, opt_kind=N.ORX_OPT_ADAGRAD, lr=B.LR, seed=1)
    g = torch.Generator(device="cpu").manual_seed(100 + rank)
    host_ids = [tuple(torch.randint(0, n, (Bsz,), generator=g, dtype=torch.int32).pin_memory() for n in (U, I, I))
                for _ in range(B.N_BATCHES)]
    dev_ids = [tuple(x.to(dev) for x in b) for b in host_ids]
    clocks = B.ClockSampler(dev.index or 0) if rank == 0 else None
    for i in range(W):
        model.step(*dev_ids[i % B.N_BATCHES])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for i in range(K):
        model.step(*dev_ids[i % B.N_BATCHES])
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
    for i in range(K):
        out = model.step(*(x.to(dev, non_blocking=True) for x in host_ids[i % B.N_BATCHES]))
        last = out.cpu()
    e1.record()
    barrier()
    t1 = time.time()
    if clocks:
        clocks.window(t0, t1)
    e2e_seconds = e0.elapsed_time(e1) * 1e-3
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
            "e2e_api": "openrec_b200.sharded.ShardedPairwise.step; pinned host ids in, global loss to host each step",
            "extra": {"last_loss": [float(x) for x in last], "total_items": I, "total_users": U}}
