"""Per-user ranking metrics -- mirrors openrec/tf2/metrics/ranking_metrics.py:8-69 on the
orx_rank_metrics kernel (one CTA per user row; integer rank counting is exact)."""
import torch

from ... import native as N
from ...tfshim.core import Tensor, convert


def _args(pos_mask, pred, excl_mask):
    return (convert(pred).t.to(torch.float32), convert(pos_mask).t, convert(excl_mask).t)


def AUC(pos_mask, pred, excl_mask):
    """[B] : #{(e,p): pred_e <= pred_p} / (n_pos * n_eval), eval = not(pos or excl) (ranking_metrics.py:8-25)."""
    p, m, x = _args(pos_mask, pred, excl_mask)
    return Tensor(N.engine().rank_metrics(p, m, x, (), want=("auc",))[0])


def NDCG(pos_mask, pred, excl_mask, at=[100]):
    """[B, len(at)] : DCG@k without ideal normaliser (ranking_metrics.py:28-47)."""
    p, m, x = _args(pos_mask, pred, excl_mask)
    return Tensor(N.engine().rank_metrics(p, m, x, tuple(at), want=("ndcg",))[1])


def Recall(pos_mask, pred, excl_mask, at=[100]):
    """[B, len(at)] (ranking_metrics.py:50-69)."""
    p, m, x = _args(pos_mask, pred, excl_mask)
    return Tensor(N.engine().rank_metrics(p, m, x, tuple(at), want=("recall",))[2])
