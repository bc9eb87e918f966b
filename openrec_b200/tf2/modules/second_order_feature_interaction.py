"""SecondOrderFeatureInteraction -- mirrors openrec/tf2/modules/second_order_feature_interaction.py:4-34.

``mode='reference'`` (default) is bug-compatible (strict-upper mask applied to a
lower-triangular matrix => zeros / diagonal only); ``mode='dlrm'`` is the intended strictly-lower
triangle of Z Z^T."""
import torch

from ...tfshim.core import Tensor, convert
from ...tfshim.keras.layers import Layer


class SecondOrderFeatureInteraction(Layer):
    def __init__(self, self_interaction=False, mode="reference"):
        super().__init__()
        if mode not in ("reference", "dlrm"):
            raise ValueError("mode must be 'reference' or 'dlrm'")
        self._self_interaction, self._mode = self_interaction, mode

    def call(self, inputs):
        from .. import mlp_ops
        feats = torch.stack([convert(x).t.to(torch.float32) for x in inputs], dim=1).contiguous()  # [B,F,D]
        return Tensor(mlp_ops.interaction_forward(feats, self._self_interaction, self._mode))
