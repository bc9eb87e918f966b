"""`import tensorflow` -> openrec_b200.tfshim (see compat/README.md)."""
import importlib
import sys

_shim = importlib.import_module("openrec_b200.tfshim")
sys.modules[__name__] = _shim
for _sub in ("keras", "keras.layers", "keras.optimizers", "keras.metrics", "keras.losses", "data"):
    sys.modules[f"{__name__}.{_sub}"] = importlib.import_module(f"openrec_b200.tfshim.{_sub}")
