"""openrec.tf2.recommenders surface (reference: openrec/tf2/recommenders/__init__.py:1-5)."""
from .bpr import BPR
from .wrmf import WRMF
from .gmf import GMF
from .ucml import UCML

__all__ = ["BPR", "WRMF", "GMF", "UCML"]
try:  # DLRM needs the MLP / interaction kernels
    from .dlrm import DLRM  # noqa: F401
    __all__.append("DLRM")
except ImportError:  # pragma: no cover
    pass


def __getattr__(name):   # row-sharded BPR / UCML (torch.distributed): imported on demand
    if name in ("ShardedBPR", "ShardedUCML"):
        from . import sharded
        return getattr(sharded, name)
    raise AttributeError(name)
