"""Dense / interaction operators behind keras Dense, MLP and SecondOrderFeatureInteraction, and the
DLRM forward/backward composition -- all arithmetic in liborx (orx_mlp_layer_*, orx_interact_*,
orx_gather_strided, orx_pred_loss); this module only sequences launches and owns activations."""
from __future__ import annotations

import torch

from .. import native as N

GEMM_KERNEL_NAME = "k_gemm_tma (TMA-fed tcgen05 kind::tf32, 3xTF32, 128x256 tile, TMEM accumulators)"


def _rows(B, n, device):
    """[B, n] activations whose row stride is a multiple of 4 floats: the TMA descriptors of the Dense-layer GEMMs need
    16-byte row strides (the 479-wide (dense_vec | interactions) block gets ld = 480); kernels take explicit strides."""
    ld = (n + 3) // 4 * 4
    t = torch.empty(B, ld, dtype=torch.float32, device=device)
    return t if ld == n else t[:, :n]


class GemmProfile:
    """bench.py's roofline hook: while active, every Dense-layer GEMM call (forward, and the dgrad + wgrad pair of a
    backward call) is bracketed by CUDA events on the launch stream; totals() -> (ms, flops, calls)."""
    active = None

    def __enter__(self):
        self.ev = []
        GemmProfile.active = self
        return self

    def __exit__(self, *a):
        GemmProfile.active = None

    def bracket(self, flops, pure=False):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.ev.append((e0, e1, flops, pure))
        return e0, e1

    def totals(self, pure_only=False):
        """(ms, flops, calls) over all bracketed Dense-layer calls, or only over the calls that are ONE tcgen05 GEMM
        launch (forward layers on the tensor-core path: bias + activation are fused into the GEMM's epilogue)."""
        torch.cuda.synchronize()
        ev = [e for e in self.ev if e[3]] if pure_only else self.ev
        return (sum(a.elapsed_time(b) for a, b, _, _ in ev), float(sum(f for _, _, f, _ in ev)), len(ev))


def _mlp_fwd(eng, x, w, b, act, y):
    p = GemmProfile.active
    if p is None:
        return eng.mlp_fwd(x, w, b, act, y)
    tc = x.shape[0] >= 64 and w.shape[0] >= 8 and w.shape[1] >= 16 and x.stride(0) % 4 == 0   # orx_launch_gemm_tc's rule
    e0, e1 = p.bracket(2.0 * x.shape[0] * w.shape[0] * w.shape[1], pure=tc)
    e0.record()
    eng.mlp_fwd(x, w, b, act, y)
    e1.record()


def _mlp_bwd(eng, x, y, w, act, dy, dx, dw, db):
    p = GemmProfile.active
    if p is None:
        return eng.mlp_bwd(x, y, w, act, dy, dx, dw, db)
    e0, e1 = p.bracket((4.0 if dx is not None else 2.0) * x.shape[0] * w.shape[0] * w.shape[1])
    e0.record()
    eng.mlp_bwd(x, y, w, act, dy, dx, dw, db)
    e1.record()

ACT = {None: 0, "linear": 0, "relu": 1, "sigmoid": 2}


def dense_forward(x, kernel, bias, activation):
    x = x.reshape(-1, x.shape[-1]).contiguous()
    y = torch.empty(x.shape[0], kernel.shape[1], dtype=torch.float32, device=x.device)
    N.engine().mlp_fwd(x, kernel, bias, ACT[activation], y)
    return y


def interaction_width(F, self_interaction):
    return F * (F + 1) // 2 if self_interaction else F * (F - 1) // 2


def interaction_forward(feats, self_interaction, mode):
    """feats [B,F,D] (last feature = the dense vector) -> [B, F(F-+1)/2]."""
    B, F, D = feats.shape
    out = torch.empty(B, interaction_width(F, self_interaction), dtype=torch.float32, device=feats.device)
    emb = feats[:, :F - 1, :].contiguous() if F > 1 else feats.new_empty(B, 0, D)
    N.engine().interact_fwd(emb, feats[:, F - 1, :].contiguous(), self_interaction, 0 if mode == "reference" else 1,
                            out)
    return out


class DLRMGraph:
    """Forward / backward of DLRM.inference + loss (recommenders/dlrm.py:63-100) on preallocated views:
    Z [B,T,D] embeddings, top_in [B, D+P] = (dense_vec | interactions) written in place by the last
    bottom layer and the interaction kernel."""

    def __init__(self, tables, bot, top, m_spa, self_interaction, mode, loss_kind, clip):
        self.tables, self.bot, self.top = tables, bot, top          # lists of tensors / (w, b, act) triples
        self.D, self.self_int, self.mode = m_spa, self_interaction, 0 if mode == "reference" else 1
        self.loss_kind, self.clip = loss_kind, clip

    def forward(self, dense, sparse, label=None, want_grad=False):
        eng = N.engine()
        dev = dense.device
        B, T, D = dense.shape[0], len(self.tables), self.D
        P = interaction_width(T + 1, self.self_int)
        c = {"dense": dense, "sparse": sparse}
        Z = c["Z"] = torch.empty(B, T, D, dtype=torch.float32, device=dev)
        for k, tab in enumerate(self.tables):                        # dlrm.py:83-85
            eng.gather_strided(tab, sparse, k, Z[:, k, :])
        top_in = c["top_in"] = _rows(B, D + P, dev)
        x, acts = dense, []
        for l, (w, b, act) in enumerate(self.bot):                   # dlrm.py:87
            last = l == len(self.bot) - 1
            if last and w.shape[1] != D:
                raise ValueError("the bottom MLP's last width must equal m_spa (tf.stack in the interaction)")
            y = top_in[:, :D] if last else torch.empty(B, w.shape[1], dtype=torch.float32, device=dev)
            _mlp_fwd(eng, x, w, b, act, y)
            acts.append(y)
            x = y
        c["bot_acts"] = acts
        eng.interact_fwd(Z, top_in[:, :D], self.self_int, self.mode, top_in[:, D:])      # dlrm.py:89-92
        x, acts = top_in, []
        for w, b, act in self.top:
            y = torch.empty(B, w.shape[1], dtype=torch.float32, device=dev)
            _mlp_fwd(eng, x, w, b, act, y)
            acts.append(y)
            x = y
        c["top_acts"] = acts
        raw = x.reshape(-1)
        pred = c["pred"] = torch.empty(B, dtype=torch.float32, device=dev)
        out4 = c["out4"] = torch.zeros(4, dtype=torch.float32, device=dev)
        lab = label if label is not None else torch.zeros(B, dtype=torch.float32, device=dev)
        c["dpred"] = torch.empty(B, dtype=torch.float32, device=dev) if want_grad else None
        eng.pred_loss(raw, lab, self.loss_kind, self.clip, pred, c["dpred"], out4)       # dlrm.py:72-73,97-98
        return c

    def backward(self, c):
        """-> (dZ [B,T,D], bottom [(dw, db)], top [(dw, db)])."""
        eng = N.engine()
        dense, top_in, Z = c["dense"], c["top_in"], c["Z"]
        B, D = dense.shape[0], self.D
        dev = dense.device
        dy = c["dpred"].reshape(B, 1)
        d_top_in = _rows(B, top_in.shape[1], dev)
        top_g = [None] * len(self.top)
        for l in range(len(self.top) - 1, -1, -1):
            w, b, act = self.top[l]
            x = top_in if l == 0 else c["top_acts"][l - 1]
            dx = d_top_in if l == 0 else torch.empty(B, w.shape[0], dtype=torch.float32, device=dev)
            dw = torch.empty_like(w)
            db = torch.empty_like(b) if b is not None else None
            _mlp_bwd(eng, x, c["top_acts"][l], w, act, dy, dx, dw, db)
            top_g[l] = (dw, db)
            dy = dx
        dZ = torch.empty_like(Z)
        eng.interact_bwd(Z, top_in[:, :D], d_top_in[:, D:], self.self_int, self.mode, dZ, d_top_in[:, :D])
        dy = d_top_in[:, :D]
        bot_g = [None] * len(self.bot)
        for l in range(len(self.bot) - 1, -1, -1):
            w, b, act = self.bot[l]
            x = dense if l == 0 else c["bot_acts"][l - 1]
            dx = None if l == 0 else torch.empty(B, w.shape[0], dtype=torch.float32, device=dev)
            dw = torch.empty_like(w)
            db = torch.empty_like(b) if b is not None else None
            _mlp_bwd(eng, x, c["bot_acts"][l], w, act, dy, dx, dw, db)
            bot_g[l] = (dw, db)
            dy = dx
        return dZ, bot_g, top_g
