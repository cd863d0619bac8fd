"""CPU checks of the C-ABI boundary: the library loads, exports every symbol include/lpb200.h
declares, and the host-only entry points (options, slack-form builder) behave like the reference."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import lp_b200
from lp_b200 import _ffi
from oracle import ipm_oracle as o
from tests.golden_problems import golden_arrays

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_if_needed():
    if not os.path.exists(_ffi.LIB_PATH):
        from lp_b200 import build
        build.build(verbose=False)


@pytest.fixture(scope="module")
def lib():
    _build_if_needed()
    return _ffi.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lpb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lpb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(lib):
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "liblpb200.so does not export %s" % n
    assert set(names) == set(_ffi.SIGNATURES), "ctypes table and header disagree"


def test_options_default_and_validation(lib):
    opt = _ffi.lpb_options()
    lib.lpb_options_default(C.byref(opt))
    # interior_point/mod.rs:51-60
    assert (opt.tol, opt.disp, opt.ip, opt.solver_type, opt.alpha0, opt.max_iter) == (1e-8, 0, 1, 0, 0.99995, 1000)
    assert lib.lpb_options_validate(C.byref(opt)) == _ffi.LPB_OK
    for bad in (0.0, 1.0, -1.0, 2.0):  # mod.rs:119-123
        opt.alpha0 = bad
        assert lib.lpb_options_validate(C.byref(opt)) == _ffi.LPB_ERR_INVALID_PARAMETER
    opt.alpha0 = 0.5
    opt.tol = 0.0  # mod.rs:124-128
    assert lib.lpb_options_validate(C.byref(opt)) == _ffi.LPB_ERR_INVALID_PARAMETER
    opt.tol = 1e-8
    opt.solver_type = _ffi.LPB_SOLVER_LEAST_SQUARES
    assert lib.lpb_options_validate(C.byref(opt)) == _ffi.LPB_ERR_UNSUPPORTED


def test_strerror_covers_every_variant(lib):
    for code in range(-5, 8):
        assert _ffi.strerror(code) != "unknown lpb status"


def test_problem_builder_matches_oracle_slack_form(lib):
    c, A_ub, b_ub, A_eq, b_eq, _, _ = golden_arrays("G1")
    pb = lp_b200.Problem.target(c).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
    ref = o.build_problem(c, A_ub, b_ub, A_eq, b_eq)
    assert np.array_equal(pb.A(), ref.A) and np.array_equal(pb.b(), ref.b) and np.array_equal(pb.c(), ref.c)
    assert pb.n_slack() == ref.n_slack and pb.c0() == 0.0
    rng = np.random.default_rng(0)
    for (mu, me, n) in [(5, 0, 7), (0, 4, 6), (3, 9, 11)]:
        c = rng.standard_normal(n)
        Au, bu = rng.standard_normal((mu, n)), rng.standard_normal(mu)
        Ae, be = rng.standard_normal((me, n)), rng.standard_normal(me)
        b = lp_b200.Problem.target(c)
        if mu:
            b = b.ub(Au, bu)
        if me:
            b = b.eq(Ae, be)
        pb = b.build()
        ref = o.build_problem(c, Au if mu else None, bu if mu else None, Ae if me else None, be if me else None)
        assert np.array_equal(pb.A(), ref.A) and np.array_equal(pb.b(), ref.b) and np.array_equal(pb.c(), ref.c)


def test_problem_builder_errors(lib):
    # linear_program.rs:134-143
    with pytest.raises(lp_b200.Unconstrained):
        lp_b200.Problem.target([1.0, 2.0]).build()
    with pytest.raises(lp_b200.IncompatibleInputDimensions):
        lp_b200.Problem.target([1.0, 2.0]).ub([[1.0, 2.0, 3.0]], [1.0]).build()
    with pytest.raises(lp_b200.IncompatibleInputDimensions):
        lp_b200.Problem.target([1.0, 2.0]).ub([[1.0, 2.0]], [1.0, 2.0]).build()
    with pytest.raises(lp_b200.IncompatibleInputDimensions):
        lp_b200.Problem.target([1.0, 2.0]).ub([[1.0, 2.0]], [1.0]).eq([[1.0]], [1.0]).build()


def test_interior_point_builder_mirrors_reference():
    # mod.rs:250-254 (default == custom().build()), :118-128 (validation)
    assert lp_b200.InteriorPoint.default() == lp_b200.InteriorPoint.custom().build()
    assert lp_b200.InteriorPoint.custom().tol(1e-6).build() != lp_b200.InteriorPoint.default()
    with pytest.raises(lp_b200.InvalidParameter):
        lp_b200.InteriorPoint.custom().alpha0(1.0).build()
    with pytest.raises(lp_b200.InvalidParameter):
        lp_b200.InteriorPoint.custom().tol(-1.0).build()
    s = (lp_b200.InteriorPoint.custom().tol(1e-8).disp(False).ip(True)
         .solver_type(lp_b200.EquationSolverType.Cholesky).alpha0(0.99995).max_iter(1000).build())
    assert s == lp_b200.InteriorPoint.default()


def test_compute_fails_loudly_without_gpu(lib):
    if lib.lpb_device_count() > 0:
        pytest.skip("a GPU is visible")
    c, A_ub, b_ub, A_eq, b_eq, _, _ = golden_arrays("G1")
    pb = lp_b200.Problem.target(c).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
    with pytest.raises(lp_b200.api.DeviceError) as e:
        lp_b200.InteriorPoint.default().solve(pb)
    assert e.value.code == _ffi.LPB_ERR_NO_DEVICE


def test_pinned_empty_degrades_to_numpy_without_a_gpu():
    import numpy as np
    import lp_b200
    a = lp_b200.pinned_empty((3, 5))
    a[:] = 2.0
    assert isinstance(a, np.ndarray) and a.shape == (3, 5) and a.dtype == np.float64 and a.sum() == 30.0
    v = np.ascontiguousarray(a, dtype=np.float64)          # what solve_batched does: a view, not a copy
    assert v.ctypes.data == a.ctypes.data
