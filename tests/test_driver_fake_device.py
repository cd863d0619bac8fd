"""Host loop + scalar logic (lp_b200/csrc/ipm_driver.hpp) against the oracle, on a TEST-ONLY CPU
stand-in for the device phase calls (tests/fake_device/fake_device.cpp).  No GPU, no liblpb200.so."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from lp_b200 import _ffi
from oracle import ipm_oracle as o
from tests.golden_problems import GOLDEN, golden_arrays

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "fake_device", "fake_device.cpp")
OUT = os.path.join(HERE, "fake_device", "_build", "libfake_ipm.so")


@pytest.fixture(scope="module")
def fake():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    deps = [SRC, os.path.join(HERE, "..", "lp_b200", "csrc", "ipm_driver.hpp"),
            os.path.join(HERE, "..", "include", "lpb200.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", SRC, "-o", OUT])
    lib = C.CDLL(OUT)
    lib.fake_solve.restype = C.c_int
    lib.fake_solve.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                               C.POINTER(_ffi.lpb_options), C.c_void_p, _ffi.c_double_p, _ffi.c_int64_p,
                               C.c_void_p, C.c_int64, _ffi.c_int64_p]
    lib.fake_options_default.argtypes = [C.POINTER(_ffi.lpb_options)]
    return lib


def run_fake(lib, pb, **kw):
    opts = _ffi.lpb_options()
    lib.fake_options_default(C.byref(opts))
    for k, v in kw.items():
        setattr(opts, k, v)
    m, n = pb.A.shape
    A = np.ascontiguousarray(pb.A)
    x = np.zeros(n)
    fun = C.c_double()
    it = C.c_int64()
    trace = np.zeros((1000, _ffi.LPB_TRACE_COLS))
    nrows = C.c_int64()
    rc = lib.fake_solve(m, n, A.ctypes.data, pb.b.ctypes.data, pb.c.ctypes.data, pb.c0, C.byref(opts),
                        x.ctypes.data, C.byref(fun), C.byref(it), trace.ctypes.data, 1000, C.byref(nrows))
    return rc, x, fun.value, it.value, trace[: nrows.value]


@pytest.mark.parametrize("name", sorted(GOLDEN))
def test_driver_reproduces_reference_known_answers(fake, name):
    c, A_ub, b_ub, A_eq, b_eq, x_ref, eps = golden_arrays(name)
    pb = o.build_problem(c, A_ub, b_ub, A_eq, b_eq)
    rc, x, fun, it, _ = run_fake(fake, pb)
    ref = o.InteriorPoint().solve(pb)
    assert rc == _ffi.LPB_OK
    assert np.abs(x[: len(x_ref)] - x_ref).max() <= eps
    assert it == ref.iteration
    assert abs(fun - ref.fun) <= 1e-8 * max(1.0, abs(ref.fun))


@pytest.mark.parametrize("ip", [1, 0])
@pytest.mark.parametrize("m,n,seed", [(16, 40, 0), (64, 128, 0), (64, 128, 1), (96, 160, 2)])
def test_driver_matches_oracle_per_iteration(fake, m, n, seed, ip):
    pb = o.build_problem(*o.synthetic_lp(m, n, seed))
    tr = []
    ref = o.InteriorPoint(ip=bool(ip)).solve(pb, trace=tr)
    rc, x, fun, it, trace = run_fake(fake, pb, ip=ip)
    assert rc == _ffi.LPB_OK
    assert abs(it - ref.iteration) <= 1
    assert np.abs(x[: len(ref.x)] - ref.x).max() < 1e-6
    assert abs(fun - ref.fun) <= 1e-8 * abs(ref.fun)
    # early iterations (well conditioned) must agree closely in every indicator
    for k in range(min(5, len(tr), len(trace))):
        r = tr[k]
        want = [r["alpha"], r["rho_p"], r["rho_d"], r["rho_A"], r["rho_g"], r["rho_mu"], r["obj"], r["bty"],
                r["tau"], r["kappa"]]
        np.testing.assert_allclose(trace[k][:10], want, rtol=1e-6, atol=1e-9)


def test_driver_status_paths(fake):
    rc, *_ = run_fake(fake, o.build_problem([1.0, 1.0], A_ub=[[1.0, 1.0]], b_ub=[-1.0]))
    assert rc == _ffi.LPB_ERR_INFEASIBLE
    rc, *_ = run_fake(fake, o.build_problem([-1.0, 0.0], A_ub=[[1.0, -1.0]], b_ub=[1.0]))
    assert rc == _ffi.LPB_ERR_UNBOUNDED
    c, A_ub, b_ub, A_eq, b_eq, _, _ = golden_arrays("G1")
    pb = o.build_problem(c, A_ub, b_ub, A_eq, b_eq)
    rc, x, _, it, _ = run_fake(fake, pb, max_iter=1)
    assert rc == _ffi.LPB_ERR_ITERATION_LIMIT_EXCEEDED and it == 1
    with pytest.raises(o.IterationLimitExceeded) as e:
        o.InteriorPoint(max_iter=1).solve(pb)
    np.testing.assert_allclose(x, e.value.x, rtol=1e-9, atol=1e-12)
    rc, *_ = run_fake(fake, pb, alpha0=1.0)
    assert rc == _ffi.LPB_ERR_INVALID_PARAMETER
    rc, *_ = run_fake(fake, pb, tol=0.0)
    assert rc == _ffi.LPB_ERR_INVALID_PARAMETER
    rc, *_ = run_fake(fake, pb, solver_type=_ffi.LPB_SOLVER_INVERSE)
    assert rc == _ffi.LPB_ERR_UNSUPPORTED


def test_numerical_problem_on_rank_deficient_rows(fake):
    # duplicated equality row -> M singular at the blind start -> pivot <= 0 (newton_equations.rs:63)
    A_eq = np.array([[1.0, 2.0, 3.0], [1.0, 2.0, 3.0]])
    pb = o.build_problem([1.0, 1.0, 1.0], A_eq=A_eq, b_eq=[1.0, 1.0])
    rc, *_ = run_fake(fake, pb)
    try:
        o.InteriorPoint(backend="scalar").solve(pb)
        ref_rc = _ffi.LPB_OK
    except o.NumericalProblem:
        ref_rc = _ffi.LPB_ERR_NUMERICAL_PROBLEM
    except o.LinearProgramError:
        ref_rc = None
    if ref_rc is not None:
        assert rc == ref_rc
