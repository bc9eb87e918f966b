"""PointwiseMSELoss -- mirrors openrec/tf2/modules/pointwise_mse_loss.py:4-31 (fused inside WRMF;
stand-alone calls reuse orx_pointwise_fwd on arange-indexed views)."""
import torch

from ... import native as N
from ...tfshim.core import Tensor, convert
from ...tfshim.keras.layers import Layer


class PointwiseMSELoss(Layer):
    def __init__(self, a=1.0, b=1.0, sigmoid=False):
        super().__init__()
        self._a, self._b, self._sigmoid = a, b, sigmoid

    def __call__(self, user_vec, item_vec, item_bias, label):
        return self.call((user_vec, item_vec, item_bias, label))

    def call(self, inputs):
        u, i, b, label = (convert(x).t.to(torch.float32) for x in inputs)
        B = u.shape[0]
        ar = torch.arange(B, dtype=torch.int32, device=u.device)
        out4 = torch.zeros(4, device=u.device)
        N.engine().pointwise_fwd(N.ORX_POINT_WRMF, N.table(u.contiguous()), N.table(i.contiguous()),
                                 N.table(b.reshape(B, 1).contiguous()), None, ar, ar, label.reshape(-1).contiguous(),
                                 out4, self._a, self._b, self._sigmoid)
        return Tensor(out4[0])
