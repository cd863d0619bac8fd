"""Host-vs-device bisection of the unrefined (p, q) solve late in the iteration.

    python tools/diag_bisect_host.py C2 16

From the GPU iterate after K iterations: the scalar S = -c.p + b.q with each of the three stages -- forming M,
factoring it, solving with the factor -- done either on the HOST (NumPy / OpenBLAS LAPACK, the arithmetic the CPU
oracle uses) or on the DEVICE (cuBLAS / cuSOLVER through torch, or liblpb200).  The right-hand side, p and the two dot
products are always evaluated on the host from the downloaded q, so only the named stages differ.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SIZES = {"C1": (512, 1024), "C2": (4096, 8192), "C3": (16384, 32768)}


def main():
    import torch
    from scipy.linalg import lapack
    import lp_b200
    from lp_b200 import _ffi
    from lp_b200.api import ResidentProblem
    from bench import synthetic_lp
    lib = _ffi.load()
    wl, K = sys.argv[1], int(sys.argv[2])
    m, n = SIZES[wl]
    c_, A_ub, b_ub, A_eq, b_eq = synthetic_lp(m, n, 0)
    pb = lp_b200.Problem.target(c_).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
    with ResidentProblem(pb) as rp:
        try:
            lp_b200.InteriorPoint.custom().max_iter(K).build().solve_resident(rp)
        except lp_b200.IterationLimitExceeded:
            pass
        x, z = rp.debug_read("x"), rp.debug_read("z")
    A, b, c = np.array(pb.A()), np.array(pb.b()), np.array(pb.c())
    d = x / z
    r = b + A.dot(d * c)

    def S_of(q):
        p = d * (A.T.dot(q) - c)
        return -c.dot(p) + b.dot(q)

    # ---- M
    M_H = A.dot(d[:, None] * A.T)
    M_H = np.triu(M_H) + np.triu(M_H, 1).T
    dA, dd = torch.from_numpy(A).cuda(), torch.from_numpy(d).cuda()
    M_Dt = (dA * dd) @ dA.T
    M_Dt = torch.triu(M_Dt) + torch.triu(M_Dt, 1).T
    M_D = M_Dt.cpu().numpy()
    print("%s iterate after %d iterations; max |M_D - M_H| / |M_H| entrywise: %.2e" % (
        wl, K, np.abs(M_D - M_H).max() / np.abs(M_H).max()))
    # liblpb200's own K1, with and without blocked accumulation (option "syrk_chain")
    M_L = {}
    for chain in (0, 1):
        hh = C.c_void_p()
        assert lib.lpb_create_bare(C.byref(hh), m, n, None) == 0
        assert lib.lpb_set_option(hh, b"syrk_chain", chain) == 0
        Mg = torch.zeros((m, m), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        assert lib.lpb_k_syrk_adat(hh, m, n, dA.data_ptr(), n, dd.data_ptr(), Mg.data_ptr(), m) == 0
        Ml = torch.tril(Mg) + torch.tril(Mg, -1).T
        M_L[chain] = Ml.cpu().numpy()
        lib.lpb_destroy(hh)
        dgH = np.sqrt(np.diag(M_H))
        print("lpb K1 (syrk_chain=%d): max |M - M_H| / sqrt(M_ii M_jj) = %.2e   (cuBLAS: %.2e)" % (
            chain, (np.abs(M_L[chain] - M_H) / np.outer(dgH, dgH)).max(), (np.abs(M_D - M_H) / np.outer(dgH, dgH)).max()))
    # ---- truth: host LAPACK + 3 refinement steps on the exact operator (host arithmetic)
    cfH, info = lapack.dpotrf(M_H, lower=1, clean=1)
    assert info == 0
    qt, _ = lapack.dpotrs(cfH, r, lower=1)
    for _ in range(3):
        p = d * (A.T.dot(qt) - c)
        dq, _ = lapack.dpotrs(cfH, b - A.dot(p), lower=1)
        qt = qt + dq
    ST = S_of(qt)
    print("truth (host, 3 refinement steps): S = %.6e" % ST)

    h = C.c_void_p()
    assert lib.lpb_create_bare(C.byref(h), m, n, None) == 0
    info_h = C.c_int32(-1)

    def factor(M, who):
        if who == "host":
            L, info = lapack.dpotrf(M, lower=1, clean=1)
            assert info == 0
            return L, None
        W = torch.from_numpy(np.ascontiguousarray(M)).cuda()
        if who == "cusolver":
            return torch.linalg.cholesky(W).cpu().numpy(), None
        torch.cuda.synchronize()
        assert lib.lpb_k_potrf(h, m, W.data_ptr(), m, C.byref(info_h)) == 0 and info_h.value == 0
        return np.tril(W.cpu().numpy()), W      # W keeps the device factor (and its inverted blocks) for lpb_k_potrs

    def solve(L, Wdev, who):
        if who == "host":
            q, info = lapack.dpotrs(L, r, lower=1)
            return q
        if who == "cusolver":
            Lt = torch.from_numpy(L).cuda()
            return torch.cholesky_solve(torch.from_numpy(r).cuda()[:, None], Lt)[:, 0].cpu().numpy()
        X = torch.from_numpy(r.copy()).cuda().reshape(1, m).contiguous()
        torch.cuda.synchronize()
        assert lib.lpb_k_potrs(h, m, Wdev.data_ptr(), m, X.data_ptr(), 1) == 0
        return X[0].cpu().numpy()

    print("%-8s %-10s %-10s %14s %10s" % ("M", "factor", "solve", "S", "rel err"))
    for Mname, M in (("host", M_H), ("cublas", M_D), ("lpb", M_L[0]), ("lpbchain", M_L[1])):
        for fname in ("host", "cusolver", "lpb"):
            L, Wdev = factor(M, fname)
            for sname in ("host", "cusolver", "lpb"):
                if sname == "lpb" and fname != "lpb":
                    continue
                S = S_of(solve(L, Wdev, sname))
                print("%-8s %-10s %-10s %14.6e %10.2e" % (Mname, fname, sname, S, abs(S - ST) / abs(ST)), flush=True)
    lib.lpb_destroy(h)


if __name__ == "__main__":
    main()
