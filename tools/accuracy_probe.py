"""Rounding-error probe of K1 (SYRK) and K2 (POTRF) against exactly rounded references.
python tools/accuracy_probe.py M N"""
import ctypes as C
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from lp_b200 import _ffi
    m, n = int(sys.argv[1]), int(sys.argv[2])
    lib = _ffi.load()
    rng = np.random.default_rng(0)
    A = rng.standard_normal((m, n))
    d = np.empty(n)
    perm = rng.permutation(n)
    d[perm[:m]] = 10 ** rng.uniform(3, 10, m)
    d[perm[m:]] = 10 ** rng.uniform(-8, -3, n - m)
    dA, dd = torch.from_numpy(A).cuda(), torch.from_numpy(d).cuda()
    dM = torch.zeros((m, m), dtype=torch.float64, device="cuda")
    h = C.c_void_p()
    assert lib.lpb_create_bare(C.byref(h), m, n, None) == 0
    assert lib.lpb_k_syrk_adat(h, m, n, dA.data_ptr(), n, dd.data_ptr(), dM.data_ptr(), m) == 0
    Mg = dM.cpu().numpy()
    Mb = A.dot(d[:, None] * A.T)
    # exact entries on a sample of the lower triangle
    eps = 2.0 ** -53
    idx = [(int(i), int(j)) for i, j in zip(rng.integers(0, m, 400), rng.integers(0, m, 400)) if j <= i][:150]
    eg, eb, es = [], [], []
    for (i, j) in idx:
        terms = A[i] * d * A[j]
        exact = math.fsum(terms)
        scale = float(np.abs(terms).sum()) * eps
        seq = 0.0
        for k0 in range(0, n, 4):  # one rounded add per 4-term group (a DMMA k4 step)
            seq += math.fsum(terms[k0:k0 + 4])
        eg.append(abs(Mg[i, j] - exact) / scale)
        eb.append(abs(Mb[i, j] - exact) / scale)
        es.append(abs(seq - exact) / scale)
    fmt = lambda v: "median %.2f  p90 %.2f  max %.2f" % (np.median(v), np.percentile(v, 90), max(v))
    print("SYRK m=%d n=%d, error in units of eps*sum|terms|:" % (m, n))
    print("  GPU DMMA kernel        ", fmt(eg))
    print("  CPU BLAS (numpy)       ", fmt(eb))
    print("  sequential k4 emulation", fmt(es))
    # POTRF backward error on the exactly symmetric BLAS matrix
    Ms = np.tril(Mb) + np.tril(Mb, -1).T
    dL = torch.from_numpy(Ms.copy()).cuda()
    info = C.c_int32(-1)
    assert lib.lpb_k_potrf(h, m, dL.data_ptr(), m, C.byref(info)) == 0 and info.value == 0
    Lg = np.tril(dL.cpu().numpy())
    Lc = np.linalg.cholesky(Ms)
    nM = np.linalg.norm(Ms)
    print("POTRF: ||L L^T - M||_F/||M||_F  GPU %.2e  LAPACK %.2e ; |Lg-Lc|/|Lc| %.2e" % (
        np.linalg.norm(Lg @ Lg.T - Ms) / nM, np.linalg.norm(Lc @ Lc.T - Ms) / nM,
        np.linalg.norm(Lg - Lc) / np.linalg.norm(Lc)))
    for key, val, name in ((b"trsm_impl", 1, "substitution TRSM"), (b"update_impl", 1, "+ plain DFMA update")):
        assert lib.lpb_set_option(h, key, val) == 0
        dL = torch.from_numpy(Ms.copy()).cuda()
        assert lib.lpb_k_potrf(h, m, dL.data_ptr(), m, C.byref(info)) == 0
        L2 = np.tril(dL.cpu().numpy())
        print("   %-22s ||L L^T - M||/||M|| %.2e" % (name, np.linalg.norm(L2 @ L2.T - Ms) / nM))
    lib.lpb_destroy(h)


if __name__ == "__main__":
    main()
