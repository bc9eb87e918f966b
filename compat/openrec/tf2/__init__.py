"""openrec.tf2 namespace alias -> openrec_b200.tf2."""
