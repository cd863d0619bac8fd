"""Root-cause probe of the factor / solve accuracy on LATE interior-point iterations.

    python tools/diag_accuracy.py C2 [C3]

For each workload: (A) iteration counts and final x for combinations of {refine, trsm_impl, solve_impl, structure}
against the committed oracle fixture; (B) per-iteration trace differences to the oracle; (C) per-stage errors on the
normal matrix of a late iterate -- SYRK vs cuBLAS, factor vs cuSOLVER (norm-wise and diagonally scaled backward
error), solves vs cholesky_solve (backward error) -- torch/cuBLAS/cuSOLVER are the CHECKERS here, never the product.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SIZES = {"C1": (512, 1024), "C2": (4096, 8192), "C3": (16384, 32768)}
COLS = ["alpha", "rho_p", "rho_d", "rho_A", "rho_g", "rho_mu", "obj", "bty", "tau", "kappa"]


def gold_for(wl):
    p = os.path.join(ROOT, "tests", "golden", "oracle_%s_seed0.json" % wl)
    g = json.load(open(p)) if os.path.exists(p) else None
    xs = {}
    if g:
        for tol, fn in [("1e-08", g.get("x_file"))] + [(k, v["x_file"]) for k, v in g.get("tighter", {}).items()]:
            if fn and os.path.exists(os.path.join(ROOT, "tests", "golden", fn)):
                xs[float(tol)] = np.load(os.path.join(ROOT, "tests", "golden", fn))
    return g, xs


def main():
    import torch
    import lp_b200
    from lp_b200 import _ffi
    from lp_b200.api import ResidentProblem
    from bench import synthetic_lp
    lib = _ffi.load()
    for wl in sys.argv[1:]:
        m, n = SIZES[wl]
        c, A_ub, b_ub, A_eq, b_eq = synthetic_lp(m, n, 0)
        pb = lp_b200.Problem.target(c).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
        del A_ub, A_eq
        gold, gx = gold_for(wl)
        git = gold["iterations"] if gold else -1
        print("==== %s %dx%d  oracle: it=%s fun=%s  x fixtures at tol %s" % (
            wl, m, n, git, gold and gold.get("fun"), sorted(gx)), flush=True)
        with ResidentProblem(pb) as rp:
            # ---------------------------------------------------------------- (A) option combinations
            combos = [dict(), dict(refine=0), dict(refine=0, update_impl=2), dict(update_impl=2),
                      dict(refine=0, update_impl=2, solve_impl=3), dict(refine=0, update_impl=2, solve_impl=2),
                      dict(refine=0, update_impl=2, trsm_impl=1, solve_impl=2), dict(refine=0, update_impl=1)]
            if os.environ.get("DIAG_ALL"):
                combos += [dict(refine=0, solve_impl=3), dict(refine=0, solve_impl=2), dict(refine=0, trsm_impl=1),
                           dict(refine=0, trsm_impl=1, solve_impl=2),
                           dict(refine=0, trsm_impl=1, solve_impl=2, structure=0), dict(solve_impl=3), dict(refine=2)]
            base = dict(refine=1, trsm_impl=0, solve_impl=0, structure=1, update_impl=0)
            traces = {}
            for combo in combos:
                opts = dict(base, **combo)
                for k, v in opts.items():
                    rp.set_option(k, v)
                for tol in ((1e-8, 1e-10) if 1e-10 in gx else (1e-8, 1e-9)):
                    t0 = time.perf_counter()
                    try:
                        res = lp_b200.InteriorPoint.custom().tol(tol).max_iter(60).build().solve_resident(rp)
                        dt = time.perf_counter() - t0
                        dx = np.abs(res.x() - gx[tol]).max() if tol in gx else float("nan")
                        print("  %-58s tol %.0e: it=%2d fun=%.12f max|dx| vs oracle %.3e  (%.2fs)" % (
                            combo or "default", tol, res.iteration(), res.fun(), dx, dt), flush=True)
                    except Exception as e:  # noqa: BLE001
                        print("  %-58s tol %.0e: %s it=%d" % (combo or "default", tol, type(e).__name__,
                                                              rp.last_iterations), flush=True)
                    if tol == 1e-8:
                        traces[json.dumps(combo, sort_keys=True)] = rp.trace().copy()
            # ---------------------------------------------------------------- (B) trace differences, default options
            if gold:
                for key in ("{}", json.dumps(dict(refine=0), sort_keys=True),
                            json.dumps(dict(refine=0, update_impl=2), sort_keys=True)):
                    tr = traces.get(key)
                    if tr is None:
                        continue
                    print("  trace vs oracle, options %s: per iteration max relative difference over %s" % (key, COLS))
                    for i in range(min(len(tr), len(gold["trace"]))):
                        g = gold["trace"][i]
                        want = np.array([g[k] for k in COLS])
                        got = tr[i][:10]
                        rel = np.abs(got - want) / np.maximum(np.abs(want), 1e-300)
                        worst = int(np.argmax(rel))
                        print("    it %2d  max rel %.2e (%s)   rho_mu rel %.2e  alpha abs %.2e" % (
                            i + 1, rel.max(), COLS[worst], rel[5], abs(got[0] - want[0])), flush=True)
            # ---------------------------------------------------------------- (C) per-stage errors on a late M
            for k, v in base.items():
                rp.set_option(k, v)
            late = max(2, (git if git > 0 else 20) - 2)
            try:
                lp_b200.InteriorPoint.custom().max_iter(late).build().solve_resident(rp)
            except lp_b200.IterationLimitExceeded:
                pass
            d = torch.from_numpy(rp.debug_read("dinv")).cuda()
            print("  late iterate (after %d iterations): Dinv spans %.2e .. %.2e" % (late, float(d.min()), float(d.max())))
        A = torch.from_numpy(pb.A()).cuda()
        bvec = torch.from_numpy(pb.b()).cuda()
        cvec = torch.from_numpy(pb.c()).cuda()
        h = C.c_void_p()
        assert lib.lpb_create_bare(C.byref(h), m, n, None) == 0
        ldm = (m + 15) // 16 * 16
        Mg = torch.zeros((m, ldm), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()  # the library runs on its own stream
        assert lib.lpb_k_syrk_adat(h, m, n, A.data_ptr(), n, d.data_ptr(), Mg.data_ptr(), ldm) == 0
        Mref = (A * d) @ A.T            # cuBLAS DGEMM: the library-grade comparator
        Mabs = (A.abs() * d) @ A.abs().T
        low = torch.tril(torch.ones((m, m), dtype=torch.bool, device="cuda"))
        e_syrk = ((Mg[:, :m] - Mref).abs() / Mabs)[low].max().item()
        print("  SYRK: max |M_gpu - M_cublas| / (|A| D |A|^T) = %.2e  (both carry ~ n eps of their own)" % e_syrk)
        del Mabs
        Msym = torch.tril(Mref) + torch.tril(Mref, -1).T
        dg = torch.sqrt(torch.diagonal(Msym))
        nM = torch.linalg.norm(Msym).item()

        def factor_errors(Lf, name):
            R = Lf @ Lf.T - Msym
            print("  POTRF %-34s ||LL^T-M||_F/||M||_F = %.2e   max |LL^T-M|_ij/sqrt(M_ii M_jj) = %.2e" % (
                name, torch.linalg.norm(R).item() / nM, (R.abs() / torch.outer(dg, dg)).max().item()), flush=True)

        Lref, info = torch.linalg.cholesky_ex(Msym)
        print("  cuSOLVER potrf info = %d" % int(info))
        factor_errors(Lref, "cuSOLVER")
        r1 = A @ (d * cvec) + bvec  # the (p, q) right-hand side of newton_equations.rs:187
        rhs2 = torch.stack([r1, torch.randn(m, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1))])

        def solve_errors(X, name):
            for k in range(2):
                res = Msym @ X[k] - rhs2[k]
                print("  SOLVE %-34s rhs %d: ||Mx-r||/(||M|| ||x|| + ||r||) = %.2e" % (
                    name, k, torch.linalg.norm(res).item() / (nM * torch.linalg.norm(X[k]).item() +
                                                               torch.linalg.norm(rhs2[k]).item())), flush=True)

        Xref = torch.cholesky_solve(rhs2.T.contiguous(), Lref).T.contiguous()
        solve_errors(Xref, "cuSOLVER potrf + potrs")
        info_h = C.c_int32(-1)
        for trsm, upd in ((0, 0), (0, 2), (1, 2), (0, 1)):
            assert lib.lpb_set_option(h, b"trsm_impl", trsm) == 0
            assert lib.lpb_set_option(h, b"update_impl", upd) == 0
            Lg = torch.zeros((m, ldm), dtype=torch.float64, device="cuda")
            Lg[:, :m] = Msym
            torch.cuda.synchronize()
            assert lib.lpb_k_potrf(h, m, Lg.data_ptr(), ldm, C.byref(info_h)) == 0
            name = "lpb trsm=%d update=%d" % (trsm, upd)
            print("  lpb potrf (%s) info = %d" % (name, info_h.value))
            Lt = torch.tril(Lg[:, :m])
            factor_errors(Lt, name)
            print("        max |L_gpu - L_cusolver| / max|L| = %.2e" % ((Lt - Lref).abs().max().item() / Lref.abs().max().item()))
            for simpl in ((0, 3, 2) if (trsm, upd) == (0, 0) else (0,)):
                assert lib.lpb_set_option(h, b"solve_impl", simpl) == 0
                Bx = rhs2.clone()
                torch.cuda.synchronize()
                rc = lib.lpb_k_potrs(h, m, Lg.data_ptr(), ldm, Bx.data_ptr(), 2)
                assert rc == 0, _ffi.last_error()
                solve_errors(Bx, "%s solve_impl=%d" % (name, simpl))
            Xm = torch.cholesky_solve(rhs2.T.contiguous(), Lt).T.contiguous()
            solve_errors(Xm, "%s factor + cuSOLVER potrs" % name)
        lib.lpb_destroy(h)
        del A, Mg, Mref, Msym
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
