"""CPU ORACLE (test infrastructure, NOT product code) for the `ripped` interior-point path.

This is a NumPy restatement of the reference's homogeneous predictor-corrector
interior-point iteration, function by function.  It exists only so that
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs can check (or time) the reference's algorithm on the
CPU.  Nothing under ``lp_b200/`` may import it.

Parity status
-------------
* Pinned on ``x`` (1e-6, and 1e-10 for G5) against every known-answer test the
  reference holds for this path (G1..G5 of SURVEY.md section 8c:
  ``src/lib.rs:84-113``, ``src/solvers/interior_point/mod.rs:181-192,256-344``,
  ``examples/symmetric.rs:10-25``) -- see ``tests/test_oracle_golden.py``.
* Intermediates (M, L, per-iteration indicators, ``fun``, ``iteration``) and the
  non-Optimal statuses are "parity unpinned": the reference has no tests for
  them and cannot be compiled in this image (no Rust toolchain), so they are
  defined by code reading only.
* The linear-algebra backend restated here is the reference's ``blas`` feature
  arm (LAPACK ``potrf('U')`` + ``potrs``, ``newton_equations.rs:85-105``);
  ``backend="scalar"`` restates the default pure-Rust arm's *shape* (lower
  Cholesky + two triangular solves, ``newton_equations.rs:129-132,151-169``)
  with an unblocked column Cholesky (linfa-linalg 0.1, not in tree).

Every function cites the reference lines it follows (paths relative to
``/root/reference/src``).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Callable, List, Optional, Tuple

import numpy as np

try:  # LAPACK through SciPy (OpenBLAS, multithreaded) -- the `blas` feature arm.
    from scipy.linalg import lapack as _lapack
except Exception:  # pragma: no cover - SciPy is present in this image
    _lapack = None


# --------------------------------------------------------------------------- errors
class LinearProgramError(Exception):
    """error.rs:7-29 -- one subclass per variant."""


class Unconstrained(LinearProgramError):
    pass


class NumericalProblem(LinearProgramError):
    pass


class InvalidParameter(LinearProgramError):
    pass


class IncompatibleInputDimensions(LinearProgramError):
    pass


class Infeasible(LinearProgramError):
    pass


class Unbounded(LinearProgramError):
    pass


class IterationLimitExceeded(LinearProgramError):
    """Carries the best x/tau in slack form (interior_point/mod.rs:237-239)."""

    def __init__(self, x):
        super().__init__("iteration limit exceeded")
        self.x = x


# --------------------------------------------------------------------------- problem
@dataclasses.dataclass
class Problem:
    """linear_program.rs:24-30 (slack form: min c'x st A x == b, x >= 0)."""

    A: np.ndarray
    b: np.ndarray
    c: np.ndarray
    c0: float
    n_slack: int

    def denormalize_target(self, x_slack: np.ndarray) -> float:
        # linear_program.rs:61-63
        return float(self.c.dot(x_slack) + self.c0)

    def denormalize_x(self, x_slack: np.ndarray) -> np.ndarray:
        # linear_program.rs:65-69
        return x_slack[: len(x_slack) - self.n_slack].copy()


def build_problem(c, A_ub=None, b_ub=None, A_eq=None, b_eq=None) -> Problem:
    """ProblemBuilder::build, linear_program.rs:125-169."""
    c = np.asarray(c, dtype=np.float64)
    n = c.shape[0]
    A_ub = np.zeros((0, n)) if A_ub is None else np.asarray(A_ub, dtype=np.float64)
    b_ub = np.zeros(0) if b_ub is None else np.asarray(b_ub, dtype=np.float64)
    A_eq = np.zeros((0, n)) if A_eq is None else np.asarray(A_eq, dtype=np.float64)
    b_eq = np.zeros(0) if b_eq is None else np.asarray(b_eq, dtype=np.float64)
    nrows_ub, ncols_ub = A_ub.shape
    nrows_eq, ncols_eq = A_eq.shape
    if nrows_ub + nrows_eq == 0:  # :134-136
        raise Unconstrained()
    if (ncols_ub != ncols_eq or ncols_eq != n or nrows_ub != b_ub.shape[0]
            or nrows_eq != b_eq.shape[0]):  # :137-143
        raise IncompatibleInputDimensions()
    A1 = np.concatenate([A_ub, A_eq], axis=0)  # :145
    A2 = np.concatenate([np.eye(nrows_ub), np.zeros((nrows_eq, nrows_ub))], axis=0)  # :147-154
    A = np.ascontiguousarray(np.concatenate([A1, A2], axis=1))  # :155
    b = np.concatenate([b_ub, b_eq])  # :157
    c_slack = np.concatenate([c, np.zeros(nrows_ub)])  # :159
    return Problem(A=A, b=b, c=c_slack, c0=0.0, n_slack=nrows_ub)  # :161-168


# --------------------------------------------------------------------------- residuals
@dataclasses.dataclass
class Residuals:
    rho_p: float
    rho_d: float
    rho_g: float
    rho_mu: float


def residuals_calculate(pb: Problem, x, y, z, tau, kappa) -> Residuals:
    """Residuals::calculate, residual.rs:13-44."""
    def norm(a):  # :22  sqrt(a.dot(a))
        return math.sqrt(float(a.dot(a)))
    rho_p = norm(pb.b * tau - pb.A.dot(x))  # :23
    rho_d = norm(pb.c * tau - pb.A.T.dot(y) - z)  # :24-26
    rho_g = abs(kappa + float(pb.c.dot(x)) - float(pb.b.dot(y)))  # :27-29,36
    rho_mu = (float(x.dot(z)) + tau * kappa) / float(len(x) + 1)  # :30-32
    return Residuals(rho_p, rho_d, rho_g, rho_mu)


# --------------------------------------------------------------------------- point / delta / rhat
@dataclasses.dataclass
class FeasiblePoint:
    """feasible_point.rs:14-21."""
    x: np.ndarray
    y: np.ndarray
    z: np.ndarray
    tau: float
    kappa: float
    initial_residuals: Residuals


@dataclasses.dataclass
class Delta:
    """delta.rs:12-18."""
    d_x: np.ndarray
    d_y: np.ndarray
    d_z: np.ndarray
    d_tau: float
    d_kappa: float


@dataclasses.dataclass
class Rhat:
    """rhat.rs:8-14."""
    p: np.ndarray
    d: np.ndarray
    g: float
    xs: np.ndarray
    tk: float


def blind_start(pb: Problem) -> FeasiblePoint:
    """FeasiblePoint::blind_start, feasible_point.rs:24-39."""
    m, n = pb.A.shape
    x = np.ones(n)
    y = np.zeros(m)
    z = np.ones(n)
    tau = 1.0
    kappa = 1.0
    return FeasiblePoint(x, y, z, tau, kappa, residuals_calculate(pb, x, y, z, tau, kappa))


def get_step_size(pt: FeasiblePoint, delta: Delta, alpha0: float) -> float:
    """feasible_point.rs:53-72 -- alpha0 multiplies AFTER the min with 1."""
    def vec_min(d, v):
        neg = d < 0.0
        if not neg.any():
            return 1.0
        return min(1.0, float(np.min(v[neg] / -d[neg])))

    def scal_min(default, d, v):
        return min(default, v / -d) if d < 0.0 else default

    alpha_x = vec_min(delta.d_x, pt.x)  # :61
    alpha_z = vec_min(delta.d_z, pt.z)  # :62
    alpha_tau = scal_min(1.0, delta.d_tau, pt.tau)  # :63
    alpha_kappa = scal_min(1.0, delta.d_kappa, pt.kappa)  # :64
    return min(min(min(min(1.0, alpha_x), alpha_tau), alpha_z), alpha_kappa) * alpha0  # :66-71


def do_step(pt: FeasiblePoint, delta: Delta, alpha: float, ip: bool) -> FeasiblePoint:
    """feasible_point.rs:76-106."""
    x = pt.x + delta.d_x * alpha
    y = pt.y + delta.d_y * alpha
    z = pt.z + delta.d_z * alpha
    tau = pt.tau + delta.d_tau * alpha
    kappa = pt.kappa + delta.d_kappa * alpha
    if ip:  # :87-95 (y is not clamped)
        x = np.maximum(x, 1.0)
        z = np.maximum(z, 1.0)
        tau = max(tau, 1.0)
        kappa = max(kappa, 1.0)
    return FeasiblePoint(x, y, z, tau, kappa, pt.initial_residuals)


def update_gamma(ip: bool, alpha: float) -> float:
    """feasible_point.rs:155-165."""
    if ip:
        return 10.0
    beta1 = 0.1
    return (1.0 - alpha) ** 2 * min(beta1, 1.0 - alpha)


def rhat_predictor(r_P, r_D, r_G, eta, pt: FeasiblePoint, gamma, mu) -> Rhat:
    """Rhat::predictor, rhat.rs:17-35."""
    return Rhat(p=r_P * eta, d=r_D * eta, g=r_G * eta,
                xs=(pt.x * -1.0) * pt.z + gamma * mu,  # :32
                tk=gamma * mu - pt.tau * pt.kappa)  # :33


def rhat_corrector(r_P, r_D, r_G, eta, pt: FeasiblePoint, delta: Delta, gamma, mu, alpha, ip) -> Rhat:
    """Rhat::corrector, rhat.rs:37-75."""
    if ip:  # :51-60
        alpha_2 = alpha * alpha
        xs = (pt.x * -1.0) * pt.z - (delta.d_x * delta.d_z) * alpha_2 + (1.0 - alpha) * gamma * mu
        tk = (1.0 - alpha) * gamma * mu - pt.tau * pt.kappa - alpha_2 * delta.d_tau * delta.d_kappa
    else:  # :62-66
        xs = (pt.x * -1.0) * pt.z + gamma * mu - (delta.d_x * delta.d_z)
        tk = gamma * mu - pt.tau * pt.kappa - delta.d_tau * delta.d_kappa
    return Rhat(p=r_P * eta, d=r_D * eta, g=r_G * eta, xs=xs, tk=tk)


# --------------------------------------------------------------------------- newton equations
class EquationsSolver:
    """EquationSolverType::build + EquationsSolver (Cholesky arm only).

    newton_equations.rs:48-64 (build), :87-90/:98-104 (blas arm) and
    :129-132/:151-169 (pure-Rust arm).  The Inverse / LeastSquares arms and the
    fallback chain (:201-209) are out of scope of the accelerated path and are
    deliberately not restated; a failed factorisation is ``NumericalProblem``
    (:63) exactly as in the reference.
    """

    def __init__(self, pt: FeasiblePoint, pb: Problem, backend: str = "lapack",
                 gemm: Optional[Callable] = None):
        self.Dinv = pt.x / pt.z  # :54
        A = pb.A
        if gemm is not None:
            self.M = gemm(A, self.Dinv)
        else:
            self.M = A.dot(self.Dinv[:, None] * A.T)  # :55-57
        self.backend = backend
        if backend == "lapack":
            # factorizec(UPLO::Upper) == LAPACK dpotrf('U')  (:88)
            c, info = _lapack.dpotrf(self.M, lower=0, clean=0, overwrite_a=0)
            if info != 0 or not np.isfinite(c.diagonal()).all():
                raise NumericalProblem()
            self.factor = c
        elif backend == "scalar":
            self.factor = _scalar_cholesky_lower(self.M)  # M.cholesky() (:130)
        else:
            raise ValueError(backend)

    def solve(self, r: np.ndarray) -> np.ndarray:
        if self.backend == "lapack":  # factor.solvec(b) == dpotrs (:100)
            v, info = _lapack.dpotrs(self.factor, r, lower=0)
            if info != 0:
                raise NumericalProblem()
            return v
        L = self.factor  # solvec_into: L w = r ; L^T v = w  (:154)
        w = _forward_sub(L, r)
        return _backward_sub_t(L, w)

    def sym_solve(self, A, r1, r2) -> Tuple[np.ndarray, np.ndarray]:
        """newton_equations.rs:214-225 ([1] eq. 8.31 / 8.32)."""
        r = r2 + A.dot(self.Dinv * r1)  # :220
        v = self.solve(r)  # :221
        u = self.Dinv * (A.T.dot(v) - r1)  # :223
        return u, v

    def solve_newton_equations(self, pb: Problem, x, rhat: Rhat):
        """newton_equations.rs:176-210 (Cholesky arm; no fallback)."""
        p, q = self.sym_solve(pb.A, pb.c, pb.b)  # :187
        u, v = self.sym_solve(pb.A, rhat.d - rhat.xs / x, rhat.p)  # :188
        if np.isnan(p).any() or np.isnan(q).any():  # :190-194
            raise NumericalProblem()
        return p, q, u, v


def _scalar_cholesky_lower(M: np.ndarray) -> np.ndarray:
    """Unblocked lower Cholesky (column by column); pivot <= 0 or non-finite is an error."""
    n = M.shape[0]
    L = np.zeros_like(M)
    for j in range(n):
        s = M[j, j] - L[j, :j].dot(L[j, :j])
        if not (s > 0.0) or not math.isfinite(s):
            raise NumericalProblem()
        d = math.sqrt(s)
        L[j, j] = d
        if j + 1 < n:
            L[j + 1:, j] = (M[j + 1:, j] - L[j + 1:, :j].dot(L[j, :j])) / d
    return L


def _forward_sub(L, r):
    n = len(r)
    w = np.empty(n)
    for i in range(n):
        w[i] = (r[i] - L[i, :i].dot(w[:i])) / L[i, i]
    return w


def _backward_sub_t(L, w):
    n = len(w)
    v = np.empty(n)
    for i in range(n - 1, -1, -1):
        v[i] = (w[i] - L[i + 1:, i].dot(v[i + 1:])) / L[i, i]
    return v


DEBUG_SCALARS: dict = {}  # scalars of the most recent Delta::compute (parity debugging; copied into the trace)


def delta_compute(pt: FeasiblePoint, rhat: Rhat, pb: Problem, solver: EquationsSolver) -> Delta:
    """Delta::compute, delta.rs:21-49."""
    p, q, u, v = solver.solve_newton_equations(pb, pt.x, rhat)  # :27
    cu, bv, cp, bq = float(pb.c.dot(u)), float(pb.b.dot(v)), float(pb.c.dot(p)), float(pb.b.dot(q))
    d_tau = ((rhat.g + 1.0 / pt.tau * rhat.tk - (-cu + bv))
             / (1.0 / pt.tau * pt.kappa + (-cp + bq)))  # :29-32
    d_x = u + p * d_tau  # :33
    d_y = v + q * d_tau  # :34
    d_z = (rhat.xs - pt.z * d_x) / pt.x  # :37
    d_kappa = 1.0 / pt.tau * (rhat.tk - pt.kappa * d_tau)  # :38
    DEBUG_SCALARS.update(cp=cp, bq=bq, cu=cu, bv=bv, d_tau=d_tau, d_kappa=d_kappa)  # last call = the corrector
    return Delta(d_x, d_y, d_z, d_tau, d_kappa)


def get_delta(pt: FeasiblePoint, pb: Problem, ip: bool, backend="lapack", gemm=None,
              timers: Optional[dict] = None) -> Delta:
    """FeasiblePoint::get_delta, feasible_point.rs:110-152."""
    n_x = len(pt.x)
    gamma = 1.0 if ip else 0.0  # :119
    eta = 1.0 if ip else 1.0 - gamma  # :120
    r_P = pb.b * pt.tau - pb.A.dot(pt.x)  # :122
    r_D = pb.c * pt.tau - pb.A.T.dot(pt.y) - pt.z  # :123
    r_G = float(pb.c.dot(pt.x)) - float(pb.b.dot(pt.y)) + pt.kappa  # :124
    mu = (float(pt.x.dot(pt.z)) + pt.tau * pt.kappa) / float(n_x + 1)  # :125

    if timers is not None:
        import time
        t0 = time.perf_counter()
    solver = EquationsSolver(pt, pb, backend=backend, gemm=gemm)  # :127
    if timers is not None:
        timers["form_factor_s"] = timers.get("form_factor_s", 0.0) + time.perf_counter() - t0

    rhat = rhat_predictor(r_P, r_D, r_G, eta, pt, gamma, mu)  # :129
    predictor_delta = delta_compute(pt, rhat, pb, solver)  # :130-131

    alpha = get_step_size(pt, predictor_delta, 1.0)  # :134
    gamma = update_gamma(ip, alpha)  # :135
    eta = 1.0 if ip else 1.0 - gamma  # :136
    rhat = rhat_corrector(r_P, r_D, r_G, eta, pt, predictor_delta, gamma, mu, alpha, ip)  # :137-148
    return delta_compute(pt, rhat, pb, solver)  # :149


# --------------------------------------------------------------------------- indicators
@dataclasses.dataclass
class Indicators:
    """indicators.rs:8-23."""
    rho_p: float
    rho_d: float
    rho_A: float
    rho_g: float
    rho_mu: float
    obj: float
    bty: float

    def display(self) -> str:
        # indicators.rs:25-33  "{:3.8}\t{:3.8}\t{:3.8}\t{:3.8}\t{:8.3}"
        return "%.8f\t%.8f\t%.8f\t%.8f\t%8.3f" % (self.rho_p, self.rho_d, self.rho_g, self.rho_mu, self.obj)

    def status(self, tau: float, kappa: float, tol: float) -> str:
        """indicators.rs:66-83 -- infeasibility test has priority over optimal."""
        tau_too_small = tau < tol * max(kappa, 1.0)
        inf1 = (self.rho_p < tol and self.rho_d < tol and self.rho_g < tol) and tau_too_small
        inf2 = self.rho_mu < tol and tau_too_small
        if inf1 or inf2:
            return "Infeasible" if self.bty > tol else "Unbounded"
        if self.rho_p < tol and self.rho_d < tol and self.rho_A < tol:
            return "Optimal"
        return "Unfinished"


def indicators_from(pt: FeasiblePoint, pb: Problem) -> Indicators:
    """Indicators::from_point_and_problem, indicators.rs:37-55."""
    obj = float(pb.c.dot(pt.x / pt.tau)) + pb.c0  # :41
    bty = float(pb.b.dot(pt.y))  # :42
    rho_A = abs(float(pb.c.dot(pt.x)) - bty) / (pt.tau + abs(float(pb.b.dot(pt.y))))  # :43-44
    res = residuals_calculate(pb, pt.x, pt.y, pt.z, pt.tau, pt.kappa)  # :45
    ini = pt.initial_residuals
    return Indicators(
        rho_p=res.rho_p / max(ini.rho_p, 1.0),  # :47
        rho_d=res.rho_d / max(ini.rho_d, 1.0),  # :48
        rho_A=rho_A,
        rho_g=res.rho_g / max(ini.rho_g, 1.0),  # :50
        rho_mu=res.rho_mu / ini.rho_mu,  # :51
        obj=obj, bty=bty)


# --------------------------------------------------------------------------- solver
@dataclasses.dataclass
class OptimizeResult:
    """solvers/mod.rs:19-49."""
    x: np.ndarray
    fun: float
    iteration: int


@dataclasses.dataclass
class InteriorPoint:
    """InteriorPointBuilder defaults + validation, interior_point/mod.rs:51-60,118-128."""
    tol: float = 1e-8
    disp: bool = False
    ip: bool = True
    solver_type: str = "Cholesky"
    alpha0: float = 0.99995
    max_iter: int = 1000
    backend: str = "lapack"

    def __post_init__(self):
        if self.alpha0 <= 0.0 or self.alpha0 >= 1.0:
            raise InvalidParameter("Alpha0 must be between 0 and 1 (exclusive)")
        if self.tol <= 0.0:
            raise InvalidParameter("The tolerance must be nonnegative.")
        if self.solver_type != "Cholesky":
            raise InvalidParameter("oracle restates the Cholesky arm only")

    def solve_normal_form(self, pb: Problem, trace: Optional[List[dict]] = None, gemm=None,
                          timers: Optional[dict] = None, stop_after: Optional[int] = None,
                          on_iteration: Optional[Callable] = None):
        """interior_point/mod.rs:199-240.  `on_iteration(iteration, point, indicators)` is a fixture-generation hook
        (tools/oracle_full_size.py snapshots x / tau where a looser tolerance would have stopped); it does not
        influence the loop."""
        pt = blind_start(pb)  # :203
        ind = indicators_from(pt, pb)  # :206
        if self.disp:  # :208-211
            print("alpha     \trho_p     \trho_d     \trho_g     \trho_mu    \tobj       ")
            print("1.00000000\t" + ind.display())
        ip = self.ip
        for iteration in range(1, self.max_iter + 1):  # :213
            delta = get_delta(pt, pb, ip, backend=self.backend, gemm=gemm, timers=timers)  # :215
            alpha = 1.0 if ip else get_step_size(pt, delta, self.alpha0)  # :216-221
            pt = do_step(pt, delta, alpha, ip)  # :222
            ip = False  # :223
            ind = indicators_from(pt, pb)  # :225
            if self.disp:  # :227-229
                print("%.8f\t%s" % (alpha, ind.display()))
            if trace is not None:
                trace.append(dict(iteration=iteration, alpha=alpha, tau=pt.tau, kappa=pt.kappa,
                                  **dataclasses.asdict(ind), **DEBUG_SCALARS))
            if on_iteration is not None:
                on_iteration(iteration, pt, ind)
            st = ind.status(pt.tau, pt.kappa, self.tol)  # :230-235
            if st == "Optimal":
                return pt.x / pt.tau, iteration
            if st == "Infeasible":
                raise Infeasible()
            if st == "Unbounded":
                raise Unbounded()
            if stop_after is not None and iteration >= stop_after:
                return pt.x / pt.tau, iteration
        raise IterationLimitExceeded(pt.x / pt.tau)  # :237-239

    def solve(self, pb: Problem, **kw) -> OptimizeResult:
        """Solver::solve, interior_point/mod.rs:161-169."""
        x_slack, iteration = self.solve_normal_form(pb, **kw)
        fun = pb.denormalize_target(x_slack)
        x = pb.denormalize_x(x_slack)
        return OptimizeResult(x, fun, iteration)


# --------------------------------------------------------------------------- synthetic inputs
def synthetic_lp(m: int, n: int, seed: int = 0):
    """SURVEY.md section 8(d) generator.  m, n are the SLACK-FORM dims.

    m_ub = m_eq = m/2, n0 = n - m/2 user variables; strictly primal- and
    dual-feasible by construction, so an optimum exists.
    Returns (c, A_ub, b_ub, A_eq, b_eq).
    """
    assert m % 2 == 0 and n > m // 2
    mh = m // 2
    n0 = n - mh
    rng = np.random.default_rng(seed)
    A0 = rng.standard_normal((m, n0))
    x0 = rng.uniform(0.5, 1.5, n0)
    s0 = rng.uniform(0.5, 1.5, mh)
    b = A0.dot(x0)
    b[:mh] += s0
    y0 = rng.standard_normal(m)
    y0[:mh] = -np.abs(y0[:mh])
    z0 = rng.uniform(0.5, 1.5, n0)
    c = A0.T.dot(y0) + z0
    return c, A0[:mh], b[:mh], A0[mh:], b[mh:]
