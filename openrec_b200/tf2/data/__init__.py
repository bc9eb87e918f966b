"""openrec.tf2.data surface (reference: openrec/tf2/data/__init__.py:1-3)."""
from .utils import _DataStore, _ParallelDataset
from .dataset import Dataset

__all__ = ["Dataset", "_DataStore", "_ParallelDataset"]


_DEVICE = ("DevicePairwiseSampler", "DeviceStratifiedSampler", "DevicePerPositiveSampler")


def __getattr__(name):   # lazy: keeps torch out of the spawned sampler workers
    if name in _DEVICE:
        from . import device_sampler
        return getattr(device_sampler, name)
    raise AttributeError(name)
