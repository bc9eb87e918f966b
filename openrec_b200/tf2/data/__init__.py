"""openrec.tf2.data surface (reference: openrec/tf2/data/__init__.py:1-3)."""
from .utils import _DataStore, _ParallelDataset
from .dataset import Dataset

__all__ = ["Dataset", "_DataStore", "_ParallelDataset"]
