"""PairwiseLogLoss -- mirrors openrec/tf2/modules/pairwise_log_loss.py:4-34.

Inside ``BPR`` the loss is fused into the training-step kernel.  Called stand-alone on materialised
vectors it runs the same forward kernel (orx_pairwise_fwd) by viewing the vectors as tiny tables
indexed by arange -- no separate arithmetic path exists."""
import torch

from ... import native as N
from ...tfshim.core import Tensor, convert
from ...tfshim.keras.layers import Layer


class PairwiseLogLoss(Layer):
    def __call__(self, user_vec, p_item_vec, n_item_vec, p_item_bias=None, n_item_bias=None):
        return self.call((user_vec, p_item_vec, n_item_vec, p_item_bias, n_item_bias))

    def call(self, inputs):
        u, p, n, bp, bn = (None if x is None else convert(x).t.to(torch.float32) for x in inputs)
        B, D = u.shape
        dev = u.device
        items = torch.cat([p, n], 0).contiguous()
        zeros = torch.zeros(B, 1, device=dev)
        bias = torch.cat([zeros if bp is None else bp.reshape(B, 1), zeros if bn is None else bn.reshape(B, 1)], 0)
        ar = torch.arange(B, dtype=torch.int32, device=dev)
        out4 = torch.zeros(4, device=dev)
        N.engine().pairwise_fwd(N.ORX_PAIR_BPR, N.table(u.contiguous()), N.table(items), N.table(bias.contiguous()),
                                ar, ar, ar + B, out4)
        return Tensor(out4[0])
