"""openrec.tf2.metrics -> openrec_b200.tf2.metrics."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("openrec_b200.tf2.metrics")
