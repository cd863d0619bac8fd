"""Which stage makes the refinement-free GPU direction noisy late in the iteration?  

    python tools/diag_dtau.py C3 17

Runs the default solve for K iterations, then -- from that SAME iterate -- evaluates the scalars of the (p, q) solve
(newton_equations.rs:187, delta.rs:29-32):   S = -c.p + b.q   (the denominator of d_tau next to kappa / tau;
mathematically p' D^-1 p >= 0, numerically a difference of two large numbers) with every combination of
{who forms M, who factors, who solves}: liblpb200 kernels vs cuBLAS / cuSOLVER (checkers only), against a
"truth" obtained by three steps of iterative refinement on the exact operator A D A^T.
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
    import lp_b200
    from lp_b200 import _ffi
    from lp_b200.api import ResidentProblem
    from bench import synthetic_lp
    lib = _ffi.load()
    wl, K = sys.argv[1], int(sys.argv[2])
    m, n = SIZES[wl]
    c_, A_ub, b_ub, A_eq, b_eq = synthetic_lp(m, n, 0)
    pb = lp_b200.Problem.target(c_).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
    del A_ub, A_eq
    with ResidentProblem(pb) as rp:
        try:
            lp_b200.InteriorPoint.custom().max_iter(K).build().solve_resident(rp)
        except lp_b200.IterationLimitExceeded:
            pass
        tr = rp.trace()
        tau, kappa = tr[-1][8], tr[-1][9]
        x = torch.from_numpy(rp.debug_read("x")).cuda()
        z = torch.from_numpy(rp.debug_read("z")).cuda()
        print("%s after %d iterations: tau %.6f kappa %.3e rho_mu %.3e" % (wl, K, tau, kappa, tr[-1][5]))
        # ---- in-library variants from this iterate (form_and_factor does not move the iterate)
        n_tot = n
        mu = float((x * z).sum().item() + tau * kappa) / (n_tot + 1)
        for opts in (dict(refine=3), dict(refine=1), dict(refine=0), dict(refine=0, syrk_chain=1), dict(refine=0, solve_impl=3),
                     dict(refine=0, solve_impl=2), dict(refine=0, solve_impl=2, trsm_impl=1),
                     dict(refine=0, update_impl=2), dict(refine=0, syrk_impl=1),
                     dict(refine=0, potf2_impl=1, solve_impl=2), dict(refine=0, potf2_impl=1, solve_impl=2, update_impl=1),
                     dict(refine=0, potf2_impl=1, solve_impl=2, update_impl=1, syrk_impl=1), dict(refine=1, refine_max=3),
                     dict(refine=2)):
            base = dict(refine=1, refine_max=0, syrk_chain=0, solve_impl=0, trsm_impl=0, structure=1, update_impl=0, syrk_impl=0, potf2_impl=0)
            base.update(opts)
            for k, v in base.items():
                rp.set_option(k, v)
            assert lib.lpb_form_and_factor(rp.handle) == 0, _ffi.last_error()
            din = _ffi.lpb_direction_in(0, 0, 1.0, 0.0, mu, 0.0)
            dout = _ffi.lpb_direction_out()
            assert lib.lpb_direction(rp.handle, C.byref(din), tau, kappa, C.byref(dout)) == 0, _ffi.last_error()
            S = -dout.cp + dout.bq
            print("  lib %-52s cp %.12e bq %.12e  S=-cp+bq %.9e  cu %.9e bv %.9e" % (
                opts, dout.cp, dout.bq, S, dout.cu, dout.bv), flush=True)
    A = torch.from_numpy(pb.A()).cuda()
    b = torch.from_numpy(pb.b()).cuda()
    c = torch.from_numpy(pb.c()).cuda()
    d = x / z
    torch.cuda.synchronize()
    r = b + A @ (d * c)                                     # newton_equations.rs:220 with (r1, r2) = (c, b)

    def scalars(q):
        p = d * (A.T @ q - c)                               # :223
        cp, bq = float(c @ p), float(b @ q)
        return cp, bq, -cp + bq, p

    def refine_truth(q, solve, steps):
        for _ in range(steps):
            p = d * (A.T @ q - c)
            res = b - A @ p                                  # residual of M q = r against the exact operator
            q = q + solve(res)
        return q

    Mc = (A * d) @ A.T
    Mc = torch.tril(Mc) + torch.tril(Mc, -1).T
    Lc = torch.linalg.cholesky(Mc)
    solve_c = lambda rhs: torch.cholesky_solve(rhs[:, None], Lc)[:, 0]
    q_a = solve_c(r)
    q_true = refine_truth(q_a.clone(), solve_c, 3)
    cpT, bqT, ST, _ = scalars(q_true)
    print("  truth (cuBLAS M, cuSOLVER factor/solve, 3 refinement steps): cp %.12e bq %.12e S %.9e" % (cpT, bqT, ST))

    def report(name, q):
        cp, bq, S, _ = scalars(q)
        print("  mix %-58s S %.9e  rel err of S %.2e   |q - q_true|/|q_true| %.2e" % (
            name, S, abs(S - ST) / abs(ST), float(torch.linalg.norm(q - q_true) / torch.linalg.norm(q_true))), flush=True)

    report("a: cuBLAS M, cuSOLVER potrf, cuSOLVER potrs", q_a)
    h = C.c_void_p()
    assert lib.lpb_create_bare(C.byref(h), m, n, None) == 0
    ldm = (m + 15) // 16 * 16
    Mg = torch.zeros((m, ldm), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    assert lib.lpb_k_syrk_adat(h, m, n, A.data_ptr(), n, d.data_ptr(), Mg.data_ptr(), ldm) == 0
    Mgs = torch.tril(Mg[:, :m]) + torch.tril(Mg[:, :m], -1).T
    Lg_c = torch.linalg.cholesky(Mgs)
    report("b: lpb SYRK M, cuSOLVER potrf, cuSOLVER potrs", torch.cholesky_solve(r[:, None], Lg_c)[:, 0])
    info = C.c_int32(-1)
    for trsm in (0, 1):
        assert lib.lpb_set_option(h, b"trsm_impl", trsm) == 0
        W = torch.zeros((m, ldm), dtype=torch.float64, device="cuda")
        W[:, :m] = Mc
        torch.cuda.synchronize()
        assert lib.lpb_k_potrf(h, m, W.data_ptr(), ldm, C.byref(info)) == 0 and info.value == 0, info.value
        Lw = torch.tril(W[:, :m])
        report("c: cuBLAS M, lpb potrf (trsm_impl=%d), cuSOLVER potrs" % trsm, torch.cholesky_solve(r[:, None], Lw)[:, 0])
        for simpl in (0, 3, 2):
            assert lib.lpb_set_option(h, b"solve_impl", simpl) == 0
            X = r.clone().reshape(1, m).contiguous()
            torch.cuda.synchronize()
            assert lib.lpb_k_potrs(h, m, W.data_ptr(), ldm, X.data_ptr(), 1) == 0, _ffi.last_error()
            report("d: cuBLAS M, lpb potrf (trsm_impl=%d), lpb potrs solve_impl=%d" % (trsm, simpl), X[0])
    # e: everything lpb, but the rhs / p / dots in torch: isolates the sweeps and the dot products of the library
    W = torch.zeros((m, ldm), dtype=torch.float64, device="cuda")
    W[:, :m] = Mgs
    assert lib.lpb_set_option(h, b"trsm_impl", 0) == 0 and lib.lpb_set_option(h, b"solve_impl", 0) == 0
    torch.cuda.synchronize()
    assert lib.lpb_k_potrf(h, m, W.data_ptr(), ldm, C.byref(info)) == 0 and info.value == 0
    X = r.clone().reshape(1, m).contiguous()
    torch.cuda.synchronize()
    assert lib.lpb_k_potrs(h, m, W.data_ptr(), ldm, X.data_ptr(), 1) == 0
    report("e: lpb SYRK + potrf + potrs, rhs and dots by torch", X[0])
    lib.lpb_destroy(h)


if __name__ == "__main__":
    main()
