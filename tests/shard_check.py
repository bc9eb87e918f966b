"""Touched-rows parity check of a row-sharded model at ANY table size (test infrastructure; uses the oracle).

One step on a fresh batch: every rank snapshots the rows (and Adagrad accumulators) of its shard that the GLOBAL batch
touches, before and after the step; rank 0 collects them, runs the float64 oracle step on that compact problem and compares
loss, l2, rows and accumulators.  Like tests/test_gpu_kernels.py::test_full_size_pairwise_adagrad, for tables that only
exist sharded (BASELINE configs[4]: 100M items over 8 GPUs).  Callers: tests/_sharded_worker.py (gloo, CPU) and
``bench.py --gpus N --check`` (through openrec_b200.sharded.bench)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


class HomeRoutedAccess:
    """Row accessors of openrec_b200.sharded.HomeRoutedPairwise (separate user / item / bias shards)."""

    def __init__(self, m):
        self.m = m

    def read(self, lu, li):
        m = self.m
        acc = lambda slots, ix: slots[0][ix].clone() if slots[0] is not None else None
        return {"user": m.user[lu].clone(), "item": m.item[li].clone(), "bias": m.bias[li].clone(),
                "user_acc": acc(m.user_slots, lu), "item_acc": acc(m.item_slots, li), "bias_acc": acc(m.bias_slots, li)}


class CombinedAccess:
    """Row accessors of openrec_b200.sharded.ShardedPairwise (one combined [user rows | item rows, D+4] table)."""

    def __init__(self, m):
        self.m = m

    def read(self, lu, li):
        m, D = self.m, self.m.D
        s0 = m.slots[0]
        it = li + m.ru
        return {"user": m.table[lu, :D].clone(), "item": m.table[it, :D].clone(), "bias": m.table[it, D:D + 1].clone(),
                "user_acc": s0[lu, :D].clone() if s0 is not None else None,
                "item_acc": s0[it, :D].clone() if s0 is not None else None,
                "bias_acc": s0[it, D:D + 1].clone() if s0 is not None else None}


def _gather_rows(t, n_max, world):
    """all_gather of a [n, w] tensor padded to n_max rows -> list of world [n_max, w] tensors (None stays None)."""
    if t is None:
        return None
    pad = torch.zeros(n_max, t.shape[1], dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return parts


def run(model, access, rank, world, total_users, total_items, dim, batch, *, kind=0, opt_kind=1, lr=0.05, margin=0.5, seed=4242,
        atol=1e-5):
    """Returns a dict (rank 0: the verdict; other ranks: {"rank": r}).  Raises AssertionError on rank 0 on a mismatch."""
    from oracle import openrec_oracle as O
    if opt_kind not in (0, 1):
        raise ValueError("the touched-rows check covers SGD and Adagrad")
    dev = model.eng.device
    g = torch.Generator(device="cpu").manual_seed(seed + rank)
    ids = [torch.randint(0, n, (batch,), generator=g, dtype=torch.int32).to(dev) for n in (total_users, total_items, total_items)]
    glob = []
    for t in ids:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        glob.append(torch.cat(parts).long())                  # the global batch, rank-major (the oracle's order)
    gu, gp, gn = glob
    mine = lambda a: torch.unique(a[a % world == rank] // world)   # sorted local rows of the touched ids this rank owns
    lu, li = mine(gu), mine(torch.cat([gp, gn]))
    pre = access.read(lu, li)
    step = model.iterations + 1
    out = model.step(*ids).detach().clone()
    post = access.read(lu, li)
    counts = torch.tensor([lu.numel(), li.numel()], dtype=torch.int64, device=dev)
    all_counts = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(all_counts, counts)
    all_counts = torch.stack(all_counts).cpu().numpy()
    mu, mi = int(all_counts[:, 0].max()), int(all_counts[:, 1].max())
    rows_u = _gather_rows(lu.reshape(-1, 1), mu, world)
    rows_i = _gather_rows(li.reshape(-1, 1), mi, world)
    shipped = {}
    for name in ("user", "user_acc"):
        for tag, src in (("pre", pre), ("post", post)):
            shipped[tag + "_" + name] = _gather_rows(src[name], mu, world)
    for name in ("item", "item_acc", "bias", "bias_acc"):
        for tag, src in (("pre", pre), ("post", post)):
            shipped[tag + "_" + name] = _gather_rows(src[name], mi, world)
    if rank != 0:
        return {"rank": rank}

    def cat(parts, col):   # the valid rows of every rank, rank after rank -> float64 numpy
        if parts is None:
            return None
        return np.concatenate([p[:all_counts[r, col]].cpu().numpy() for r, p in enumerate(parts)]).astype(np.float64)

    gid_u = np.concatenate([rows_u[r][:all_counts[r, 0], 0].cpu().numpy() * world + r for r in range(world)])
    gid_i = np.concatenate([rows_i[r][:all_counts[r, 1], 0].cpu().numpy() * world + r for r in range(world)])

    def compact(gids, x):  # position of global id x in the concatenated row arrays
        order = np.argsort(gids, kind="stable")
        pos = np.searchsorted(gids[order], x)
        assert np.array_equal(gids[order][pos], x), "a touched row is missing from the shards' snapshots"
        return order[pos].astype(np.int64)

    cu, cp, cn = compact(gid_u, gu.cpu().numpy()), compact(gid_i, gp.cpu().numpy()), compact(gid_i, gn.cpu().numpy())
    user, item, bias = cat(shipped["pre_user"], 0), cat(shipped["pre_item"], 1), cat(shipped["pre_bias"], 1)
    st = {"user": (cat(shipped["pre_user_acc"], 0), None), "item": (cat(shipped["pre_item_acc"], 1), None),
          "bias": (cat(shipped["pre_bias_acc"], 1), None)}
    loss, l2 = O.pairwise_train_step("bpr" if kind == 0 else "ucml", user, item, bias, cu, cp, cn,
                                     O.OPT_ADAGRAD if opt_kind == 1 else O.OPT_SGD, st, step, lr, margin=margin)
    got = out.cpu().numpy()
    np.testing.assert_allclose(got[:2], [loss, l2], rtol=3e-5, err_msg="global (loss, l2)")
    worst = 0.0
    for name, ref, col in (("user", user, 0), ("item", item, 1), ("bias", bias, 1)):
        a = cat(shipped["post_" + name], col)
        np.testing.assert_allclose(a, ref, atol=atol, err_msg=name + " rows after the step")
        worst = max(worst, float(np.abs(a - ref).max()))
        if opt_kind == 1:
            np.testing.assert_allclose(cat(shipped["post_" + name + "_acc"], col), st[name][0], atol=atol,
                                       err_msg=name + " accumulators after the step")
    return {"passed": True, "global_batch": int(batch * world), "touched_user_rows": int(len(gid_u)),
            "touched_item_rows": int(len(gid_i)), "max_abs_row_error": worst, "loss": [float(got[0]), float(got[1])],
            "oracle_loss": [float(loss), float(l2)], "tolerance": atol,
            "what": "one extra step on a fresh batch; rows + Adagrad accumulators of every shard that the GLOBAL batch touches, "
                    "before and after, against the float64 oracle step on the compacted problem (tests/shard_check.py)"}
