"""openrec.tf2.modules surface (reference: openrec/tf2/modules/__init__.py:1-5)."""
from .latent_factor import LatentFactor
from .pairwise_log_loss import PairwiseLogLoss
from .pointwise_mse_loss import PointwiseMSELoss
from .multi_layer_perceptron import MLP
from .second_order_feature_interaction import SecondOrderFeatureInteraction

__all__ = ["LatentFactor", "PairwiseLogLoss", "PointwiseMSELoss", "MLP", "SecondOrderFeatureInteraction"]
