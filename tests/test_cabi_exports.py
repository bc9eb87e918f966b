"""CPU-only: liborx.so loads, exports exactly the symbols include/orx.h declares, and refuses to
compute without a device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from openrec_b200 import build
    return build.build()


def _declared():
    src = open(os.path.join(ROOT, "include", "orx.h")).read()
    return sorted(set(re.findall(r"ORX_API\s+(?:const\s+char\*|int)\s+(orx_\w+)\s*\(", src)))


def test_header_symbols_exported(built):
    out = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\sT\s+(orx_\w+)", out)))
    assert _declared() == exported


def test_ctypes_signatures_cover_header(built):
    from openrec_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    l = _lib.lib()
    assert l.orx_abi_version() == 1


def test_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "--list-elf", built], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("device present")
    from openrec_b200 import _lib, native
    h = C.c_void_p()
    rc = _lib.lib().orx_create(0, C.byref(h))
    assert rc != 0 and not h.value
    assert _lib.last_error()
    with pytest.raises(RuntimeError):
        native.engine()
