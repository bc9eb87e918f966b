"""openrec namespace alias -> openrec_b200 (see compat/README.md)."""
