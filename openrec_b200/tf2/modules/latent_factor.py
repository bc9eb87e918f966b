"""LatentFactor -- mirrors openrec/tf2/modules/latent_factor.py:4-23 on liborx."""
from ... import native as N
from ...tfshim.core import convert
from ...tfshim.keras.layers import Embedding


class LatentFactor(Embedding):
    """Embedding table [num_instances, dim]; U(-0.05, 0.05) init unless zero_init (latent_factor.py:6-15).
    Calling it gathers rows (orx_gather); ``variables[0]`` is the table (bpr.py:42)."""

    def __init__(self, num_instances, dim, zero_init=False, name=None):
        super().__init__(input_dim=num_instances, output_dim=dim,
                         embeddings_initializer="zeros" if zero_init else "uniform", name=name)

    def censor(self, censor_id):
        """rows of the unique ids <- row / max(||row||, 0.1)  (latent_factor.py:17-23), in place."""
        N.engine().censor(self.embeddings.t, convert(censor_id).t, 0.1)
        return self.embeddings
