"""openrec.tf2.modules -> openrec_b200.tf2.modules."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("openrec_b200.tf2.modules")
