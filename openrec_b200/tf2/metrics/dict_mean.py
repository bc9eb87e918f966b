"""DictMean -- mirrors openrec/tf2/metrics/dict_mean.py:4-32 (running per-key mean over axis 0)."""
import torch

from ...tfshim.core import Tensor, convert, device


class DictMean:
    def __init__(self, state_shape):
        self._shapes = {k: tuple(int(s) for s in shp) for k, shp in state_shape.items()}
        self.reset_states()

    def reset_states(self):
        self._sum = {k: torch.zeros(shp, dtype=torch.float32, device=device()) for k, shp in self._shapes.items()}
        self._count = {k: 0.0 for k in self._shapes}

    def update_state(self, state):
        for k, v in state.items():
            t = convert(v).t.to(torch.float32)
            self._sum[k] += t.sum(dim=0)
            self._count[k] += float(t.shape[0])

    def result(self):
        return {k: Tensor(self._sum[k] / self._count[k]) if self._count[k] else Tensor(self._sum[k] * float("nan"))
                for k in self._sum}
