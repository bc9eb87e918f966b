"""ctypes front of oracle/c/orx_oracle.c (CPU baseline arm + cross-check of the numpy oracle)."""
import ctypes as C
import os

import numpy as np

from . import build_c

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build_c.LIB if os.path.exists(build_c.LIB) else build_c.build()
        l = C.CDLL(path)
        l.orx_oracle_pairwise_step.restype = C.c_int
        l.orx_oracle_num_threads.restype = C.c_int
        _lib = l
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def num_threads():
    return lib().orx_oracle_num_threads()


def pairwise_step(kind, user, user_acc, item, item_acc, bias, bias_acc, uid, pid, nid, opt, lr, margin=0.5,
                  c_loss=1.0, c_l2=1.0, eps=1e-7, nthreads=0):
    """In-place float32 step; kind 'bpr'|'ucml'; opt 0 SGD | 1 Adagrad.  Returns (loss, l2_loss)."""
    for a in (user, item, bias):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    out = (C.c_double * 2)()
    rc = lib().orx_oracle_pairwise_step(
        C.c_int(0 if kind == "bpr" else 1), _p(user), _p(user_acc), _p(item), _p(item_acc), _p(bias), _p(bias_acc),
        C.c_int64(user.shape[0]), C.c_int64(item.shape[0]), C.c_int(user.shape[1]), _p(uid), _p(pid), _p(nid),
        C.c_int(len(uid)), C.c_float(margin), C.c_float(c_loss), C.c_float(c_l2), C.c_int(opt), C.c_float(lr),
        C.c_float(eps), C.c_int(nthreads), out)
    if rc != 0:
        raise ValueError("orx_oracle_pairwise_step: bad arguments / out-of-range ids")
    return out[0], out[1]
