"""tf.keras: Model / Sequential + the submodules the examples import."""
from . import layers, losses, metrics, optimizers  # noqa: F401
from .layers import Layer


class Model(Layer):
    """tf.keras.Model: a Layer whose __call__ forwards positional and keyword arguments to call()."""

    def __init__(self, *args, **kwargs):
        super().__init__(name=kwargs.get("name"))


class Sequential(Model):
    def __init__(self, layers=None, name=None):
        super().__init__(name=name)
        self.layers = list(layers or [])

    def add(self, layer):
        self.layers = self.layers + [layer]   # reassign: invalidates cached variable lists

    def build(self, in_dim):
        for l in self.layers:
            l.build(in_dim)
            in_dim = l.units

    def call(self, x):
        for l in self.layers:
            x = l(x)
        return x
