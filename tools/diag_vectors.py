"""Pinpoint where the GPU path loses accuracy against LAPACK late in the iteration: run K iterations on
the GPU through the phase calls, then compare every intermediate of the next predictor solve with a CPU
recomputation from the SAME iterate.   python tools/diag_vectors.py M N K [key=value ...]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import scipy.linalg as sl
    import lp_b200
    from lp_b200 import _ffi
    from lp_b200.api import ResidentProblem
    from bench import synthetic_lp
    m, n, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    opts = dict(kv.split("=") for kv in sys.argv[4:])
    lib = _ffi.load()
    cc, A_ub, b_ub, A_eq, b_eq = synthetic_lp(m, n, 0)
    pb = lp_b200.Problem.target(cc).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
    A, b, c = pb.A(), pb.b(), pb.c()
    with ResidentProblem(pb) as rp:
        for k, v in opts.items():
            rp.set_option(k, int(v))
        h = rp.handle
        tau = kappa = 1.0
        assert lib.lpb_blind_start(h) == 0
        rs = _ffi.lpb_residual_scalars()
        assert lib.lpb_residuals(h, tau, kappa, C.byref(rs)) == 0
        ip = True
        for it in range(1, K + 2):
            gamma = 1.0 if ip else 0.0
            eta = 1.0 if ip else 1.0 - gamma
            r_G = rs.cx - rs.by + kappa
            mu = (rs.xz + tau * kappa) / (n + 1)
            if it == K + 1:
                x, z, y = rp.debug_read("x"), rp.debug_read("z"), rp.debug_read("y")
            assert lib.lpb_form_and_factor(h) == 0
            din = _ffi.lpb_direction_in(0, int(ip), eta, gamma, mu, 0.0)
            dout = _ffi.lpb_direction_out()
            assert lib.lpb_direction(h, C.byref(din), tau, kappa, C.byref(dout)) == 0
            if it == K + 1:
                break
            tk = gamma * mu - tau * kappa

            def dscal(g_hat, tk):
                d_tau = (g_hat + 1.0 / tau * tk - (-dout.cu + dout.bv)) / (1.0 / tau * kappa + (-dout.cp + dout.bq))
                return d_tau, 1.0 / tau * (tk - kappa * d_tau)

            def step(axz, d_tau, d_kappa, a0):
                at = min(1.0, tau / -d_tau) if d_tau < 0 else 1.0
                ak = min(1.0, kappa / -d_kappa) if d_kappa < 0 else 1.0
                return min(1.0, axz[0], at, axz[1], ak) * a0

            d_tau, d_kappa = dscal(r_G * eta, tk)
            axz = (C.c_double * 2)()
            assert lib.lpb_assemble_delta(h, d_tau, axz) == 0
            alpha = step(axz, d_tau, d_kappa, 1.0)
            one_m = 1.0 - alpha
            gamma = 10.0 if ip else (one_m * one_m) * min(0.1, one_m)
            eta = 1.0 if ip else 1.0 - gamma
            tk = ((1.0 - alpha) * gamma * mu - tau * kappa - alpha * alpha * d_tau * d_kappa) if ip else (
                gamma * mu - tau * kappa - d_tau * d_kappa)
            din = _ffi.lpb_direction_in(1, int(ip), eta, gamma, mu, alpha)
            assert lib.lpb_direction(h, C.byref(din), tau, kappa, C.byref(dout)) == 0
            d_tau, d_kappa = dscal(r_G * eta, tk)
            assert lib.lpb_assemble_delta(h, d_tau, axz) == 0
            alpha = 1.0 if ip else step(axz, d_tau, d_kappa, 0.99995)
            assert lib.lpb_do_step(h, alpha, int(ip)) == 0
            tau, kappa = tau + d_tau * alpha, kappa + d_kappa * alpha
            if ip:
                tau, kappa = max(tau, 1.0), max(kappa, 1.0)
            ip = False
            assert lib.lpb_residuals(h, tau, kappa, C.byref(rs)) == 0
        # ---- GPU intermediates of the predictor solve at iteration K+1
        Lg = np.tril(rp.debug_read("M").reshape(m, -1)[:, :m])
        W = rp.debug_read("W")
        v_g, q_g = W[:m], W[m:]
        t = rp.debug_read("t")
        p_g, u_g, dinv_g = rp.debug_read("p")[:n], rp.debug_read("u")[:n], rp.debug_read("dinv")[:n]
        cp_g, bq_g = dout.cp, dout.bq
    print("iteration %d: tau=%.6f kappa=%.3e  GPU cp=%.12f bq=%.12f bq-cp=%.4e" % (K + 1, tau, kappa, cp_g, bq_g, bq_g - cp_g))
    Dinv = x / z
    print("dinv max rel diff gpu vs x/z: %.2e ; range %.2e..%.2e" % (np.abs(dinv_g / Dinv - 1).max(), Dinv.min(), Dinv.max()))
    M = A.dot(Dinv[:, None] * A.T)
    Lc = np.linalg.cholesky(M)
    print("L: ||Lg - Lc||_F/||Lc||_F = %.2e ; ||Lg Lg^T - M||_F/||M||_F = %.2e (LAPACK %.2e)" % (
        np.linalg.norm(Lg - Lc) / np.linalg.norm(Lc), np.linalg.norm(Lg @ Lg.T - M) / np.linalg.norm(M),
        np.linalg.norm(Lc @ Lc.T - M) / np.linalg.norm(M)))
    s = A.dot(Dinv * c)
    print("A(Dinv c): ||t1_gpu - s||/||s|| = %.2e" % (np.linalg.norm(t[m:2 * m] - s) / np.linalg.norm(s)))
    r = b + s
    r_g = b + t[m:2 * m]
    q_c = sl.cho_solve((Lc, True), r)
    q_mixL = sl.cho_solve((Lg, True), r_g)          # GPU factor + GPU rhs, CPU substitution
    pDp = lambda p: float((p * p / Dinv).sum())
    for name, q in (("CPU (LAPACK)", q_c), ("GPU L + CPU substitution", q_mixL), ("GPU", q_g)):
        p = Dinv * (A.T.dot(q) - c)
        res = r - M @ q
        print("%-26s bq=%.12f cp(CPU gemv)=%.12f bq-cp=%.4e pDp=%.4e |q-q_c|/|q|=%.2e s.(q-q_c)=%.3e q.res=%.3e" % (
            name, b.dot(q), c.dot(p), b.dot(q) - c.dot(p), pDp(p), np.linalg.norm(q - q_c) / np.linalg.norm(q_c),
            s.dot(q - q_c), q.dot(res)))
    p_from_qg = Dinv * (A.T.dot(q_g) - c)
    print("p: ||p_gpu - Dinv(A^T q_gpu - c)||/||p|| = %.2e ; c.p_gpu=%.12f vs CPU-gemv c.p=%.12f" % (
        np.linalg.norm(p_g - p_from_qg) / np.linalg.norm(p_from_qg), c.dot(p_g), c.dot(p_from_qg)))
    atq = A.T.dot(q_g)
    atq_g = p_g / Dinv + c
    big = np.argsort(-Dinv)[:5]
    print("largest-Dinv columns: Dinv=%s  (A^T q - c) CPU=%s GPU=%s" % (Dinv[big], (atq - c)[big], (atq_g - c)[big]))


if __name__ == "__main__":
    main()
