"""openrec.tf2.recommenders -> openrec_b200.tf2.recommenders."""
import importlib
import sys

sys.modules[__name__] = importlib.import_module("openrec_b200.tf2.recommenders")
