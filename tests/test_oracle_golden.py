"""Pin the CPU oracle against every known-answer test the reference holds for the IPM path."""
import numpy as np
import pytest

from oracle import ipm_oracle as o
from tests.golden_problems import GOLDEN, golden_arrays, symmetric_example


@pytest.mark.parametrize("name", sorted(GOLDEN))
@pytest.mark.parametrize("backend", ["lapack", "scalar"])
def test_reference_known_answers(name, backend):
    c, A_ub, b_ub, A_eq, b_eq, x_ref, eps = golden_arrays(name)
    pb = o.build_problem(c, A_ub, b_ub, A_eq, b_eq)
    res = o.InteriorPoint(backend=backend).solve(pb)
    assert np.abs(res.x - x_ref).max() <= eps
    assert res.iteration >= 1


def test_symmetric_example_G5():
    c, A_ub, b_ub, _, _, x_ref, eps = symmetric_example(1000)
    pb = o.build_problem(c, A_ub, b_ub)
    res = o.InteriorPoint().solve(pb)
    assert np.abs(res.x - x_ref).max() <= eps          # examples/symmetric.rs:21
    assert abs(res.fun - (-1000.0)) < 1e-6


def test_default_equals_custom_G6():
    # interior_point/mod.rs:250-254
    assert o.InteriorPoint() == o.InteriorPoint(tol=1e-8, disp=False, ip=True, alpha0=0.99995, max_iter=1000)


def test_builder_validation():
    # interior_point/mod.rs:118-128
    for bad in (0.0, 1.0, -0.5, 1.5):
        with pytest.raises(o.InvalidParameter):
            o.InteriorPoint(alpha0=bad)
    with pytest.raises(o.InvalidParameter):
        o.InteriorPoint(tol=0.0)


def test_problem_builder_errors():
    # linear_program.rs:134-143
    with pytest.raises(o.Unconstrained):
        o.build_problem([1.0, 2.0])
    with pytest.raises(o.IncompatibleInputDimensions):
        o.build_problem([1.0, 2.0], A_ub=[[1.0, 2.0, 3.0]], b_ub=[1.0])
    with pytest.raises(o.IncompatibleInputDimensions):
        o.build_problem([1.0, 2.0], A_ub=[[1.0, 2.0]], b_ub=[1.0, 2.0])


def test_slack_form_layout():
    # linear_program.rs:145-161: A = [[A_ub, I], [A_eq, 0]], b = [b_ub; b_eq], c = [c; 0]
    c, A_ub, b_ub, A_eq, b_eq, _, _ = golden_arrays("G1")
    pb = o.build_problem(c, A_ub, b_ub, A_eq, b_eq)
    assert pb.A.shape == (3, 4) and pb.n_slack == 2
    assert np.array_equal(pb.A, np.array([[-3, 1, 1, 0], [1, 2, 0, 1], [1, 1, 0, 0]], dtype=float))
    assert np.array_equal(pb.b, [6, 4, 1]) and np.array_equal(pb.c, [-1, 4, 0, 0])


def test_status_paths_infeasible_unbounded():
    # not covered by the reference's tests; defined by code reading of indicators.rs:66-83
    with pytest.raises(o.Infeasible):   # x1 + x2 <= -1, x >= 0
        o.InteriorPoint().solve(o.build_problem([1.0, 1.0], A_ub=[[1.0, 1.0]], b_ub=[-1.0]))
    with pytest.raises(o.Unbounded):    # min -x1 st x1 - x2 <= 1
        o.InteriorPoint().solve(o.build_problem([-1.0, 0.0], A_ub=[[1.0, -1.0]], b_ub=[1.0]))


def test_iteration_limit_carries_x():
    c, A_ub, b_ub, A_eq, b_eq, _, _ = golden_arrays("G1")
    pb = o.build_problem(c, A_ub, b_ub, A_eq, b_eq)
    with pytest.raises(o.IterationLimitExceeded) as e:
        o.InteriorPoint(max_iter=1).solve(pb)
    assert e.value.x.shape == (4,)


@pytest.mark.parametrize("m,n,iters", [(64, 128, 9), (256, 512, None)])
def test_synthetic_generator_solves(m, n, iters):
    pb = o.build_problem(*o.synthetic_lp(m, n, seed=0))
    assert pb.A.shape == (m, n)
    res = o.InteriorPoint().solve(pb)
    res2 = o.InteriorPoint(backend="scalar").solve(pb) if m <= 64 else res
    assert abs(res.iteration - res2.iteration) <= 1
    assert np.abs(res.x - res2.x).max() < 1e-6
    if iters is not None:
        assert res.iteration == iters


def test_scipy_ancestor_agrees_on_goldens():
    """SciPy's _ip_hsd (the algorithm's ancestor) as a second, non-bit-exact cross-check."""
    _ip = pytest.importorskip("scipy.optimize._linprog_ip")
    for name in sorted(GOLDEN):
        c, A_ub, b_ub, A_eq, b_eq, x_ref, eps = golden_arrays(name)
        pb = o.build_problem(c, A_ub, b_ub, A_eq, b_eq)
        x, status, *_ = _ip._ip_hsd(pb.A, pb.b, pb.c, 0.0, alpha0=0.99995, beta=0.1, maxiter=1000,
                                    disp=False, tol=1e-8, sparse=False, lstsq=False, sym_pos=True,
                                    cholesky=True, pc=True, ip=True, permc_spec="MMD_AT_PLUS_A",
                                    callback=None, postsolve_args=None)
        assert status == 0
        assert np.abs(x[: len(x_ref)] - x_ref).max() < 1e-6
