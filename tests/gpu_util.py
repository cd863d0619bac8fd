"""Helpers for the -m gpu tests: device arrays via torch (plumbing only), calls through the C ABI."""
import ctypes as C

import numpy as np

from lp_b200 import _ffi


def torch_mod():
    import torch
    return torch


def to_dev(a):
    torch = torch_mod()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def pad_cols(a, mult=16):
    """Row-major copy whose leading dimension is a multiple of `mult` doubles."""
    a = np.asarray(a, dtype=np.float64)
    ld = (a.shape[1] + mult - 1) // mult * mult
    out = np.zeros((a.shape[0], ld))
    out[:, : a.shape[1]] = a
    return out, ld


class BareCtx:
    def __init__(self, m, n):
        self.lib = _ffi.load()
        self.h = C.c_void_p()
        rc = self.lib.lpb_create_bare(C.byref(self.h), m, n, None)
        assert rc == 0, _ffi.last_error()

    def set(self, key, val):
        assert self.lib.lpb_set_option(self.h, key.encode(), val) == 0

    def close(self):
        if self.h:
            self.lib.lpb_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def ok(rc):
    assert rc == 0, "lpb rc=%d: %s" % (rc, _ffi.last_error())
