"""Where does the host-side time of lpb_solve_batched go?  (diagnostic; GPU box only)"""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

import lp_b200
from lp_b200 import _ffi

lib = _ffi.load()
batch, m, n = 8192, 64, 128
rng = np.random.default_rng(0)
for kind in ("pageable", "pinned"):
    if kind == "pinned":
        A, b, c = lp_b200.pinned_empty((batch, m, n)), lp_b200.pinned_empty((batch, m)), lp_b200.pinned_empty((batch, n))
    else:
        A, b, c = np.empty((batch, m, n)), np.empty((batch, m)), np.empty((batch, n))
    A0 = rng.standard_normal((m, n - m // 2))
    for i in range(batch):
        A[i, :, : n - m // 2] = A0
        A[i, :, n - m // 2:] = 0.0
        A[i, np.arange(m // 2), n - m // 2 + np.arange(m // 2)] = 1.0
    x0 = rng.uniform(0.5, 1.5, n)
    b[:] = A[0] @ x0
    c[:] = A[0].T @ rng.standard_normal(m) + rng.uniform(0.5, 1.5, n)
    for rep in range(3):
        t0 = time.perf_counter()
        r = lp_b200.solve_batched(A, b, c, n_slack=m // 2)
        t1 = time.perf_counter()
        print(kind, "solve_batched call %d: %.1f ms (optimal %d)" % (rep, (t1 - t0) * 1e3, int((r.status == 0).sum())), flush=True)
    import torch
    t = torch.empty(A.size, dtype=torch.float64, device="cuda")
    src = torch.from_numpy(np.asarray(A).reshape(-1))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    print(kind, "plain H2D of A: %.1f ms" % ((time.perf_counter() - t0) * 1e3), flush=True)
