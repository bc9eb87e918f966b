"""Thin torch-tensor front of the C-ABI: torch owns device memory and streams (plumbing),
liborx does all the arithmetic.  Every function enqueues on torch's current CUDA stream."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import (ORX_OPT_ADAGRAD, ORX_OPT_ADAM_DENSE, ORX_OPT_ADAM_LAZY, ORX_OPT_SGD, ORX_PAIR_BPR,
                   ORX_PAIR_UCML, ORX_POINT_GMF, ORX_POINT_WRMF, ORX_SCORE_DOT, ORX_SCORE_NEG_SQDIST, OrxOpt,
                   OrxTable)

__all__ = ["Engine", "engine", "table", "opt", "ORX_PAIR_BPR", "ORX_PAIR_UCML", "ORX_POINT_GMF", "ORX_POINT_WRMF",
           "ORX_OPT_SGD", "ORX_OPT_ADAGRAD", "ORX_OPT_ADAM_LAZY", "ORX_OPT_ADAM_DENSE", "ORX_SCORE_DOT",
           "ORX_SCORE_NEG_SQDIST"]

_engines = {}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _f32(t, name):
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise ValueError(f"{name}: expected a contiguous float32 CUDA tensor")
    return t


def ids32(t):
    """int32 contiguous CUDA ids (Keras Embedding casts other integer dtypes to int32)."""
    if not t.is_cuda:
        raise ValueError("ids must live on the CUDA device")
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t.contiguous().reshape(-1)


def table(var, s0=None, s1=None):
    """orx_table_t for a [rows, dim] variable and its optimizer slots."""
    _f32(var, "var"), _f32(s0, "s0"), _f32(s1, "s1")
    rows, dim = (var.shape[0], var.shape[1]) if var.dim() == 2 else (1, var.numel())
    return OrxTable(var.data_ptr(), s0.data_ptr() if s0 is not None else None,
                    s1.data_ptr() if s1 is not None else None, rows, dim)


def opt(kind, lr, eps=1e-7, beta1=0.9, beta2=0.999, step=1):
    return OrxOpt(kind, lr, eps, beta1, beta2, step)


class Engine:
    """One liborx context (workspace) per CUDA device."""

    def __init__(self, index: int):
        self.index = index
        self.lib = _lib.lib()
        h = C.c_void_p()
        _lib.check(self.lib.orx_create(index, C.byref(h)), "orx_create")
        self.h = h
        self.device = torch.device("cuda", index)

    def close(self):
        if self.h:
            self.lib.orx_destroy(self.h)
            self.h = None

    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    make_table = staticmethod(table)
    make_opt = staticmethod(opt)

    # ---- measurement hook ----
    def profile_enable(self, on=True):
        _lib.check(self.lib.orx_profile_enable(self.h, 1 if on else 0))

    def profile_read(self, n_phases=3):
        """-> ([ms per phase] summed over the recorded steps, n_steps); phases: see orx_profile_enable in orx.h."""
        ms = (C.c_float * n_phases)()
        n = C.c_int32()
        _lib.check(self.lib.orx_profile_read(self.h, ms, n_phases, C.byref(n)))
        return list(ms), n.value

    # ---- LatentFactor ------------------------------------------------------------------
    def fill_uniform(self, dst, lo, hi, seed):
        _lib.check(self.lib.orx_fill_uniform(self.h, _ptr(_f32(dst, "dst")), dst.numel(), lo, hi, seed, self.stream()))

    def gather(self, tab, ids, n_bad=None):
        is64 = ids.dtype == torch.int64
        if not is64:
            ids = ids32(ids)
        ids = ids.contiguous().reshape(-1)
        out = torch.empty((ids.numel(), tab.shape[1]), dtype=torch.float32, device=tab.device)
        _lib.check(self.lib.orx_gather(self.h, _ptr(_f32(tab, "tab")), tab.shape[0], tab.shape[1], _ptr(ids),
                                       1 if is64 else 0, ids.numel(), _ptr(out), _ptr(n_bad), self.stream()))
        return out

    def censor(self, tab, ids, min_norm=0.1):
        ids = ids32(ids)
        _lib.check(self.lib.orx_censor(self.h, _ptr(_f32(tab, "tab")), tab.shape[0], tab.shape[1], _ptr(ids),
                                       ids.numel(), min_norm, self.stream()))

    # ---- pairwise ----------------------------------------------------------------------
    def pairwise_step(self, kind, user, item, bias, uid, pid, nid, o, out4, margin=0.5, c_loss=1.0, c_l2=1.0):
        _lib.check(self.lib.orx_pairwise_step(self.h, kind, C.byref(user), C.byref(item), C.byref(bias), _ptr(uid),
                                              _ptr(pid), _ptr(nid), uid.numel(), margin, c_loss, c_l2, C.byref(o),
                                              _ptr(out4), self.stream()), "orx_pairwise_step")

    def pairwise_step_host(self, kind, user, item, bias, uid_h, pid_h, nid_h, o, out4_h, margin=0.5, c_loss=1.0,
                           c_l2=1.0):
        """ids / out4 are (pinned) HOST tensors; copies ride the same stream as the kernels."""
        _lib.check(self.lib.orx_pairwise_step_host(self.h, kind, C.byref(user), C.byref(item), C.byref(bias),
                                                   _ptr(uid_h), _ptr(pid_h), _ptr(nid_h), uid_h.numel(), margin,
                                                   c_loss, c_l2, C.byref(o), _ptr(out4_h), self.stream()),
                   "orx_pairwise_step_host")

    def pairwise_prefetch(self, user, item, uid, pid, nid, opt_kind, ids_ready=False):
        """Pipelining hint: build the batch index of these id tensors on the side stream now (the next pairwise_step
        with these very tensors consumes it).  ids_ready=True: the tensors are already complete (pre-staged batches),
        so the build does not wait for anything queued on the current stream."""
        _lib.check(self.lib.orx_pairwise_prefetch(self.h, C.byref(user), C.byref(item), _ptr(uid), _ptr(pid), _ptr(nid),
                                                  uid.numel(), opt_kind, 1 if ids_ready else 0, self.stream()),
                   "orx_pairwise_prefetch")

    def debug_set_epoch(self, epoch):
        _lib.check(self.lib.orx_debug_set_epoch(self.h, epoch), "orx_debug_set_epoch")

    def pairwise_fwd(self, kind, user, item, bias, uid, pid, nid, out4, margin=0.5):
        _lib.check(self.lib.orx_pairwise_fwd(self.h, kind, C.byref(user), C.byref(item), C.byref(bias), _ptr(uid),
                                             _ptr(pid), _ptr(nid), uid.numel(), margin, _ptr(out4), self.stream()),
                   "orx_pairwise_fwd")

    def pairwise_grad(self, kind, user, item, bias, uid, pid, nid, margin=0.5, c_loss=1.0, c_l2=1.0, *, d_user=None,
                      d_pos=None, d_neg=None, d_bp=None, d_bn=None, g_out=None):
        _lib.check(self.lib.orx_pairwise_grad(self.h, kind, C.byref(user), C.byref(item), C.byref(bias), _ptr(uid),
                                              _ptr(pid), _ptr(nid), uid.numel(), margin, c_loss, c_l2, _ptr(d_user),
                                              _ptr(d_pos), _ptr(d_neg), _ptr(d_bp), _ptr(d_bn), _ptr(g_out),
                                              self.stream()), "orx_pairwise_grad")

    def pairwise_grad_slots(self, kind, user_rows, item_rows, bias_rows, uslot, pslot, nslot, inv_B, d_user, d_item,
                            d_bias, out4, margin=0.5, c_loss=1.0, c_l2=1.0):
        _lib.check(self.lib.orx_pairwise_grad_slots(self.h, kind, _ptr(user_rows), _ptr(item_rows), _ptr(bias_rows),
                                                    user_rows.shape[1], _ptr(uslot), _ptr(pslot), _ptr(nslot),
                                                    uslot.numel(), margin, c_loss, c_l2, inv_B, _ptr(d_user),
                                                    _ptr(d_item), _ptr(d_bias), _ptr(out4), self.stream()),
                   "orx_pairwise_grad_slots")

    # ---- un-fused sparse apply / multi-GPU building blocks ------------------------------
    def sparse_apply(self, tab, ids, values, o):
        n = 0 if ids is None else ids.numel()
        _lib.check(self.lib.orx_sparse_apply(self.h, C.byref(tab), _ptr(ids), _ptr(values), n, C.byref(o),
                                             self.stream()), "orx_sparse_apply")

    def owner_bucket(self, ids, world):
        """-> (counts[world], send_local[n], slot[n]) device int32 tensors."""
        n = ids.numel()
        counts = torch.empty(world, dtype=torch.int32, device=ids.device)
        send_local = torch.empty(n, dtype=torch.int32, device=ids.device)
        slot = torch.empty(n, dtype=torch.int32, device=ids.device)
        _lib.check(self.lib.orx_owner_bucket(self.h, _ptr(ids), n, world, _ptr(counts), _ptr(send_local), _ptr(slot),
                                             self.stream()), "orx_owner_bucket")
        return counts, send_local, slot

    def sparse_apply_strided(self, tab, ids2d, col, values3d, o):
        """ids = ids2d[:, col] (int32 [n, F]); value rows = values3d[:, col, :] ([n, F, D]) -- no copies."""
        n, F = ids2d.shape
        D = values3d.shape[2]
        _lib.check(self.lib.orx_sparse_apply_strided(
            self.h, C.byref(tab), C.c_void_p(ids2d.data_ptr() + 4 * col), F,
            C.c_void_p(values3d.data_ptr() + 4 * col * D), values3d.shape[1] * D, n, C.byref(o), self.stream()),
            "orx_sparse_apply_strided")

    # ---- DLRM pieces (2-D operands may be column-slices: the leading dimension is taken from stride(0)) ----
    @staticmethod
    def _ld(t):
        if t.dim() != 2 or t.stride(1) != 1:
            raise ValueError("expected a 2-D float32 view with unit inner stride")
        return t.stride(0)

    def gather_strided(self, tab, ids2d, col, out2d):
        n, F = ids2d.shape
        _lib.check(self.lib.orx_gather_strided(self.h, _ptr(tab), tab.shape[0], tab.shape[1],
                                               C.c_void_p(ids2d.data_ptr() + 4 * col), F, n, _ptr(out2d),
                                               self._ld(out2d), None, self.stream()), "orx_gather_strided")

    def mlp_fwd(self, x, w, bias, act, y):
        _lib.check(self.lib.orx_mlp_layer_fwd(self.h, _ptr(x), self._ld(x), x.shape[0], w.shape[0], _ptr(w), _ptr(bias),
                                              w.shape[1], act, _ptr(y), self._ld(y), self.stream()),
                   "orx_mlp_layer_fwd")

    def mlp_bwd(self, x, y, w, act, dy, dx, dw, db):
        _lib.check(self.lib.orx_mlp_layer_bwd(self.h, _ptr(x), self._ld(x), _ptr(y), self._ld(y), _ptr(w), x.shape[0],
                                              w.shape[0], w.shape[1], act, _ptr(dy), self._ld(dy), _ptr(dx),
                                              self._ld(dx) if dx is not None else 0, _ptr(dw), _ptr(db),
                                              self.stream()), "orx_mlp_layer_bwd")

    def interact_fwd(self, emb3d, dense2d, self_interaction, mode, out2d):
        B, Fm1, D = emb3d.shape
        _lib.check(self.lib.orx_interact_fwd(self.h, _ptr(emb3d), Fm1 * D, _ptr(dense2d), self._ld(dense2d), B,
                                             Fm1 + 1, D, int(self_interaction), mode, _ptr(out2d), self._ld(out2d),
                                             self.stream()), "orx_interact_fwd")

    def interact_bwd(self, emb3d, dense2d, dout2d, self_interaction, mode, demb3d, ddense2d):
        B, Fm1, D = emb3d.shape
        _lib.check(self.lib.orx_interact_bwd(self.h, _ptr(emb3d), Fm1 * D, _ptr(dense2d), self._ld(dense2d),
                                             _ptr(dout2d), self._ld(dout2d), B, Fm1 + 1, D, int(self_interaction),
                                             mode, _ptr(demb3d), Fm1 * D, _ptr(ddense2d), self._ld(ddense2d),
                                             self.stream()), "orx_interact_bwd")

    def pred_loss(self, pred, label, kind, clip, pred_out, dpred, out4):
        _lib.check(self.lib.orx_pred_loss(self.h, _ptr(pred), _ptr(label), pred.numel(), kind, clip, _ptr(pred_out),
                                          _ptr(dpred), _ptr(out4), self.stream()), "orx_pred_loss")

    def owner_bucket_combined(self, ids, n_user, total_users, world):
        """ids = uid | pid | nid -> (counts[world], send_local[n] combined local rows, slot[n])."""
        n = ids.numel()
        counts = torch.empty(world, dtype=torch.int32, device=ids.device)
        send_local = torch.empty(n, dtype=torch.int32, device=ids.device)
        slot = torch.empty(n, dtype=torch.int32, device=ids.device)
        _lib.check(self.lib.orx_owner_bucket_combined(self.h, _ptr(ids), n, n_user, total_users, world, _ptr(counts),
                                                      _ptr(send_local), _ptr(slot), self.stream()),
                   "orx_owner_bucket_combined")
        return counts, send_local, slot

    def pairwise_grad_rows(self, kind, rows, dim, uslot, pslot, nslot, inv_B, d_rows, out4, margin=0.5, c_loss=1.0,
                           c_l2=1.0):
        _lib.check(self.lib.orx_pairwise_grad_rows(self.h, kind, _ptr(rows), rows.shape[1], dim, _ptr(uslot),
                                                   _ptr(pslot), _ptr(nslot), uslot.numel(), margin, c_loss, c_l2,
                                                   inv_B, _ptr(d_rows), _ptr(out4), self.stream()),
                   "orx_pairwise_grad_rows")

    # ---- pointwise ---------------------------------------------------------------------
    def pointwise_step(self, kind, user, item, bias, w, uid, iid, label, o, out4, a=1.0, b=1.0, use_sigmoid=False,
                       c_loss=1.0, c_l2=1.0):
        _lib.check(self.lib.orx_pointwise_step(self.h, kind, C.byref(user), C.byref(item), C.byref(bias),
                                               C.byref(w) if w is not None else None, _ptr(uid), _ptr(iid),
                                               _ptr(label), uid.numel(), a, b, int(use_sigmoid), c_loss, c_l2,
                                               C.byref(o), _ptr(out4), self.stream()), "orx_pointwise_step")

    def pointwise_fwd(self, kind, user, item, bias, w, uid, iid, label, out4, a=1.0, b=1.0, use_sigmoid=False):
        _lib.check(self.lib.orx_pointwise_fwd(self.h, kind, C.byref(user), C.byref(item), C.byref(bias),
                                              C.byref(w) if w is not None else None, _ptr(uid), _ptr(iid),
                                              _ptr(label), uid.numel(), a, b, int(use_sigmoid), _ptr(out4),
                                              self.stream()), "orx_pointwise_fwd")

    def pointwise_grad(self, kind, user, item, bias, w, uid, iid, label, a=1.0, b=1.0, use_sigmoid=False, c_loss=1.0,
                       c_l2=1.0, *, d_user=None, d_item=None, d_bias=None, d_w=None, g_out=None):
        _lib.check(self.lib.orx_pointwise_grad(self.h, kind, C.byref(user), C.byref(item), C.byref(bias),
                                               C.byref(w) if w is not None else None, _ptr(uid), _ptr(iid),
                                               _ptr(label), uid.numel(), a, b, int(use_sigmoid), c_loss, c_l2,
                                               _ptr(d_user), _ptr(d_item), _ptr(d_bias), _ptr(d_w), _ptr(g_out),
                                               self.stream()), "orx_pointwise_grad")

    # ---- dense / inference / metrics -----------------------------------------------------
    def dense_apply(self, var, s0, s1, grad, o):
        _lib.check(self.lib.orx_dense_apply(self.h, _ptr(_f32(var, "var")), _ptr(s0), _ptr(s1),
                                            _ptr(_f32(grad, "grad")), var.numel(), C.byref(o), self.stream()),
                   "orx_dense_apply")

    def score_all(self, kind, user_tab, uid, item_tab, item_bias, scale=None):
        uid = ids32(uid)
        out = torch.empty((uid.numel(), item_tab.shape[0]), dtype=torch.float32, device=item_tab.device)
        _lib.check(self.lib.orx_score_all(self.h, kind, _ptr(_f32(user_tab, "user_tab")), user_tab.shape[0], _ptr(uid),
                                          uid.numel(), _ptr(scale), _ptr(_f32(item_tab, "item_tab")),
                                          _ptr(item_bias), item_tab.shape[0], item_tab.shape[1], _ptr(out),
                                          self.stream()), "orx_score_all")
        return out

    def rank_metrics(self, pred, pos, excl, at=(), want=("auc", "ndcg", "recall")):
        pred = _f32(pred.contiguous(), "pred")
        pos = pos.to(torch.uint8).contiguous()
        excl = excl.to(torch.uint8).contiguous()
        R, I = pred.shape
        at_arr = (C.c_int32 * max(len(at), 1))(*[int(k) for k in at])
        auc = torch.empty(R, dtype=torch.float32, device=pred.device) if "auc" in want else None
        ndcg = torch.empty((R, len(at)), dtype=torch.float32, device=pred.device) if "ndcg" in want else None
        rec = torch.empty((R, len(at)), dtype=torch.float32, device=pred.device) if "recall" in want else None
        _lib.check(self.lib.orx_rank_metrics(self.h, _ptr(pred), _ptr(pos), _ptr(excl), R, I, at_arr, len(at),
                                             _ptr(auc), _ptr(ndcg), _ptr(rec), self.stream()), "orx_rank_metrics")
        return auc, ndcg, rec


def engine(device=None) -> Engine:
    """The per-device engine; raises (no CPU fallback) when CUDA is unavailable."""
    if not torch.cuda.is_available():
        raise RuntimeError("openrec_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    idx = torch.cuda.current_device() if device is None else torch.device(device).index or 0
    e = _engines.get(idx)
    if e is None:
        e = _engines[idx] = Engine(idx)
    return e
