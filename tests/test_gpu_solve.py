"""Whole-solve parity on the B200 through the reference-shaped API (Problem / InteriorPoint) and the
C ABI, against the CPU oracle.  Bar (BASELINE.json north_star): same termination status, iteration
count within +-1, x within 1e-6 absolute, objective within 1e-8 relative."""
import ctypes as C

import numpy as np
import pytest

import lp_b200
from lp_b200 import _ffi
from lp_b200.api import ResidentProblem
from oracle import ipm_oracle as o
from tests.golden_problems import GOLDEN, golden_arrays, symmetric_example

pytestmark = pytest.mark.gpu


def build(c, A_ub, b_ub, A_eq, b_eq):
    b = lp_b200.Problem.target(c)
    if A_ub is not None:
        b = b.ub(A_ub, b_ub)
    if A_eq is not None:
        b = b.eq(A_eq, b_eq)
    return b.build()


def assert_parity(res, ref, x_tol=1e-6, f_tol=1e-8):
    assert abs(res.iteration() - ref.iteration) <= 1
    assert np.abs(res.x() - ref.x).max() <= x_tol
    assert abs(res.fun() - ref.fun) <= f_tol * max(1.0, abs(ref.fun))


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_reference_known_answers(name):
    """The reference's own unit tests / doctests (G1..G4), assert_abs_diff_eq!(x, expected, 1e-6)."""
    c, A_ub, b_ub, A_eq, b_eq, x_ref, eps = golden_arrays(name)
    res = lp_b200.InteriorPoint.default().solve(build(c, A_ub, b_ub, A_eq, b_eq))
    assert np.abs(res.x() - x_ref).max() <= eps
    ref = o.InteriorPoint().solve(o.build_problem(c, A_ub, b_ub, A_eq, b_eq))
    assert res.iteration() == ref.iteration
    assert_parity(res, ref)


def test_symmetric_example_G5():
    """examples/symmetric.rs: N=1000, x == 1 to 1e-10."""
    c, A_ub, b_ub, _, _, x_ref, eps = symmetric_example(1000)
    res = lp_b200.InteriorPoint.custom().disp(True).build().solve(build(c, A_ub, b_ub, None, None))
    assert np.abs(res.x() - x_ref).max() <= eps
    assert abs(res.fun() + 1000.0) < 1e-6
    assert res.iteration() == 4


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("ip", [True, False])
@pytest.mark.parametrize("m,n,seed", [(64, 128, 0), (64, 128, 3), (96, 200, 1), (130, 301, 2), (512, 1024, 0)])
def test_synthetic_parity_with_oracle(m, n, seed, ip, impl):
    """SURVEY 8(d) generator incl. config C1 (512x1024) and ragged sizes; per-iteration trace agrees."""
    if (m % 2) or n <= m // 2:
        pytest.skip("generator needs even m")
    args = o.synthetic_lp(m, n, seed)
    tr = []
    ref = o.InteriorPoint(ip=ip).solve(o.build_problem(*args), trace=tr)
    pb = build(*args)
    solver = lp_b200.InteriorPoint.custom().ip(ip).build()
    with ResidentProblem(pb) as rp:
        rp.set_option("syrk_impl", impl)
        res = solver.solve_resident(rp)
        trace = rp.trace()
        prof = rp.profile()
    assert_parity(res, ref)
    assert prof["launches"] > 0 and prof["iterations"] == res.iteration()
    assert len(trace) == res.iteration()
    for k in range(min(len(tr), len(trace))):  # every iteration; tolerance as in test_gpu_accuracy (grows like 1 / mu)
        r = tr[k]
        want = np.array([r["alpha"], r["rho_p"], r["rho_d"], r["rho_A"], r["rho_g"], r["rho_mu"], r["obj"], r["bty"],
                         r["tau"], r["kappa"]])
        rtol = 1e-9 + 3e-8 / r["rho_mu"]
        if rtol <= 0.5:
            np.testing.assert_allclose(trace[k][:10], want, rtol=rtol, atol=1e-9)


def test_status_paths():
    """Infeasible / Unbounded / IterationLimitExceeded / InvalidParameter, vs the oracle's outcome."""
    with pytest.raises(lp_b200.Infeasible):
        lp_b200.InteriorPoint.default().solve(build([1.0, 1.0], [[1.0, 1.0]], [-1.0], None, None))
    with pytest.raises(lp_b200.Unbounded):
        lp_b200.InteriorPoint.default().solve(build([-1.0, 0.0], [[1.0, -1.0]], [1.0], None, None))
    c, A_ub, b_ub, A_eq, b_eq, _, _ = golden_arrays("G1")
    pb = build(c, A_ub, b_ub, A_eq, b_eq)
    with pytest.raises(lp_b200.IterationLimitExceeded) as e:
        lp_b200.InteriorPoint.custom().max_iter(1).build().solve(pb)
    with pytest.raises(o.IterationLimitExceeded) as eo:
        o.InteriorPoint(max_iter=1).solve(o.build_problem(c, A_ub, b_ub, A_eq, b_eq))
    np.testing.assert_allclose(e.value.x, eo.value.x, rtol=1e-9, atol=1e-12)  # slack-form x/tau, mod.rs:237-239
    with pytest.raises(lp_b200.InvalidParameter):
        lp_b200.InteriorPoint.custom().solver_type(lp_b200.EquationSolverType.Inverse).build().solve(pb)


def test_numerical_problem_on_non_finite_normal_matrix():
    """A factorisation that cannot succeed (M overflows to inf) is NumericalProblem on the GPU exactly as in
    the reference (newton_equations.rs:63: pivot not > 0 / not finite), for both oracle backends."""
    A_ub = np.array([[1e200, 1.0], [1.0, 2.0]])
    for backend in ("lapack", "scalar"):
        with pytest.raises(o.NumericalProblem):
            o.InteriorPoint(backend=backend).solve(o.build_problem([1.0, 1.0], A_ub=A_ub, b_ub=[1.0, 1.0]))
    with pytest.raises(lp_b200.NumericalProblem):
        lp_b200.InteriorPoint.default().solve(build([1.0, 1.0], A_ub, [1.0, 1.0], None, None))


def test_exactly_singular_normal_matrix_is_roundoff_dependent():
    """Duplicated equality rows make M exactly singular; whether the second pivot lands at 0, -1ulp or
    +1ulp depends on summation order and FMA contraction, so the GPU may either report NumericalProblem
    (what both oracle backends happen to do) or push through -- in which case the answer must be the
    optimum of the LP with the redundant row dropped."""
    A_eq = np.array([[1.0, 2.0, 3.0], [1.0, 2.0, 3.0]])
    ref = o.InteriorPoint().solve(o.build_problem([1.0, 1.0, 1.0], A_eq=A_eq[:1], b_eq=[1.0]))
    try:
        res = lp_b200.InteriorPoint.default().solve(build([1.0, 1.0, 1.0], None, None, A_eq, [1.0, 1.0]))
    except lp_b200.NumericalProblem:
        return
    assert np.abs(res.x() - ref.x).max() < 1e-6
    assert abs(res.fun() - ref.fun) < 1e-8


def test_host_driven_phase_calls_equal_lpb_solve():
    """The exported phase calls (what the Rust shim would drive) reproduce lpb_solve exactly."""
    lib = _ffi.load()
    args = o.synthetic_lp(64, 128, 5)
    pb = build(*args)
    res = lp_b200.InteriorPoint.default().solve(pb)
    with ResidentProblem(pb) as rp:
        h = rp.handle
        n = rp.n
        tau = kappa = 1.0
        assert lib.lpb_blind_start(h) == 0
        rs = _ffi.lpb_residual_scalars()
        assert lib.lpb_residuals(h, tau, kappa, C.byref(rs)) == 0
        ini = (rs.nrm_rp, rs.nrm_rd, abs(kappa + rs.cx - rs.by), (rs.xz + tau * kappa) / (n + 1))
        ip = True
        it = 0
        while True:
            it += 1
            gamma = 1.0 if ip else 0.0
            eta = 1.0 if ip else 1.0 - gamma
            r_G = rs.cx - rs.by + kappa
            mu = (rs.xz + tau * kappa) / (n + 1)
            assert lib.lpb_form_and_factor(h) == 0
            din = _ffi.lpb_direction_in(0, int(ip), eta, gamma, mu, 0.0)
            dout = _ffi.lpb_direction_out()
            assert lib.lpb_direction(h, C.byref(din), tau, kappa, C.byref(dout)) == 0
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
            one_m = 1.0 - alpha  # (one_m * one_m), not pow(): libm's pow is not correctly rounded
            gamma = 10.0 if ip else (one_m * one_m) * min(0.1, one_m)
            eta = 1.0 if ip else 1.0 - gamma
            if ip:
                tk = (1.0 - alpha) * gamma * mu - tau * kappa - alpha * alpha * d_tau * d_kappa
            else:
                tk = gamma * mu - tau * kappa - d_tau * d_kappa
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
            rho_p = rs.nrm_rp / max(ini[0], 1.0)
            rho_d = rs.nrm_rd / max(ini[1], 1.0)
            rho_A = abs(rs.cx - rs.by) / (tau + abs(rs.by))
            if rho_p < 1e-8 and rho_d < 1e-8 and rho_A < 1e-8:
                break
            assert it < 50
        x = np.zeros(n)
        fun = C.c_double()
        assert lib.lpb_extract_x(h, tau, x.ctypes.data, C.byref(fun)) == 0
    assert it == res.iteration()
    np.testing.assert_array_equal(x[: len(res.x())], res.x())
    assert fun.value == res.fun()


def _factor_on_device(pb, structure):
    """form_and_factor at the blind start; returns (lower factor L, lpb_profile.syrk_cols after a solve)."""
    with ResidentProblem(pb) as rp:
        rp.set_option("structure", structure)
        lib = _ffi.load()
        assert lib.lpb_blind_start(rp.handle) == 0
        assert lib.lpb_form_and_factor(rp.handle) == 0
        m = rp.m
        ldm = (m + 15) // 16 * 16
        L = np.tril(rp.debug_read("M").reshape(m, ldm)[:, :m])
        res = lp_b200.InteriorPoint.default().solve_resident(rp)
        return L, rp.profile()["syrk_cols"], res


@pytest.mark.parametrize("m,n,seed", [(64, 128, 0), (130, 301, 2), (512, 1024, 0), (300, 1500, 5)])
def test_slack_columns_are_folded_into_the_diagonal(m, n, seed):
    """The slack block [I; 0] of linear_program.rs:145-156 is not contracted over: the SYRK runs on the
    n - n_slack dense columns and d_slack lands on the diagonal.  Same factor (1e-12 relative), same solve."""
    args = o.synthetic_lp(m, n, seed)
    pb = build(*args)
    L1, cols1, res1 = _factor_on_device(pb, 1)
    L0, cols0, res0 = _factor_on_device(pb, 0)
    assert cols0 == n and cols1 == n - pb.n_slack()
    assert np.abs(L1 - L0).max() <= 1e-12 * np.abs(L0).max()
    assert res1.iteration() == res0.iteration()
    assert np.abs(res1.x() - res0.x()).max() <= 1e-7   # two roundings of the same iteration, stopped at tol = 1e-8
    assert abs(res1.fun() - res0.fun()) <= 1e-10 * max(1.0, abs(res0.fun()))


def test_structure_detection_edge_cases():
    """Scaled singleton columns fold as s^2 d; two singleton columns hitting the same row stop the run
    (the earlier one stays in the dense part); a short run (< 16 columns) is left dense."""
    rng = np.random.default_rng(11)
    m, n0 = 40, 60
    A0 = rng.standard_normal((m, n0))
    tail = np.zeros((m, 24))
    for j in range(24):
        tail[j, j] = 1.0 + 0.25 * j          # scaled unit columns, distinct rows 0..23
    tail[:, 3] = 0.0                          # an all-zero column inside the run
    x0 = rng.uniform(0.5, 1.5, n0 + 24)
    A = np.hstack([A0, tail])
    b = A @ x0
    c = A.T @ rng.standard_normal(m) + rng.uniform(0.5, 1.5, n0 + 24)
    lib = _ffi.load()

    def factor(Amat, structure):
        bufs = [lp_b200.api._HostBuffer(s) for s in (Amat.shape, b.shape, (Amat.shape[1],))]
        bufs[0].array[:] = Amat
        bufs[1].array[:] = b
        bufs[2].array[:] = c[: Amat.shape[1]]
        pb = lp_b200.Problem(bufs[0], bufs[1], bufs[2], 0.0, 0)
        with ResidentProblem(pb) as rp:
            rp.set_option("structure", structure)
            assert lib.lpb_blind_start(rp.handle) == 0
            assert lib.lpb_form_and_factor(rp.handle) == 0
            ldm = (m + 15) // 16 * 16
            L = np.tril(rp.debug_read("M").reshape(m, ldm)[:, :m])
            try:
                lp_b200.InteriorPoint.custom().max_iter(1).build().solve_resident(rp)
            except lp_b200.LinearProgramError:
                pass
            return L, rp.profile()["syrk_cols"]

    Lref = np.linalg.cholesky(A @ A.T)        # blind start: x = z = 1, Dinv = 1
    L, cols = factor(A, 1)
    assert cols == n0
    assert np.abs(L - Lref).max() <= 1e-11 * np.abs(Lref).max()
    # collision: the last column repeats row 5 -> only the columns behind the first repeat are folded
    A2 = A.copy()
    A2[:, -1] = 0.0
    A2[5, -1] = 3.0
    L2, cols2 = factor(A2, 1)
    Lref2 = np.linalg.cholesky(A2 @ A2.T)
    assert cols2 == n0 + 6                    # columns n0+6 .. end fold (row 5 is taken by the last one)
    assert np.abs(L2 - Lref2).max() <= 1e-11 * np.abs(Lref2).max()
    # odd number of dense columns (the TMA box runs past the tensor map's last column: zero fill)
    A4 = np.hstack([A0[:, :59], tail])
    L4, cols4 = factor(A4, 1)
    assert cols4 == 59
    assert np.abs(L4 - np.linalg.cholesky(A4 @ A4.T)).max() <= 1e-11 * np.abs(Lref).max()
    # short run: 8 singleton columns only -> dense
    A3 = A[:, : n0 + 8]
    L3, cols3 = factor(A3, 1)
    assert cols3 == n0 + 8
    assert np.abs(L3 - np.linalg.cholesky(A3 @ A3.T)).max() <= 1e-11 * np.abs(Lref).max()
