"""openrec.tf2.data -> openrec_b200.tf2.data."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("openrec_b200.tf2.data")
