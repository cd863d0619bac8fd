"""Opt-in presolve (lp_b200.presolve / lpb_presolve_*): host-only, so the reduction itself is
checked on the CPU -- the oracle solves the original and the presolved problem -- and once more on the GPU."""
import numpy as np
import pytest

import lp_b200
from oracle import ipm_oracle as o


def _messy_lp(seed=0):
    """A feasible, bounded LP with two duplicated rows (one scaled by -3), one empty row and rows / columns whose
    scales differ by 1e6."""
    rng = np.random.default_rng(seed)
    m0, n0 = 12, 30
    A0 = rng.standard_normal((m0, n0))
    x0 = rng.uniform(0.5, 1.5, n0)
    y0 = rng.standard_normal(m0)
    c = A0.T @ y0 + rng.uniform(0.5, 1.5, n0)
    rs = 10.0 ** rng.integers(-3, 4, m0)
    cs = 10.0 ** rng.integers(-3, 4, n0)
    As = rs[:, None] * A0 * cs[None, :]          # same LP in the variables x / cs
    bs = As @ (x0 / cs)
    A_eq = np.vstack([As, As[3], -3.0 * As[7], np.zeros(n0)])
    b_eq = np.concatenate([bs, [bs[3], -3.0 * bs[7], 0.0]])
    return c * cs, A_eq, b_eq, m0


def test_presolve_drops_redundant_rows_and_equilibrates():
    c, A_eq, b_eq, m0 = _messy_lp()
    pb = lp_b200.Problem.target(c).eq(A_eq, b_eq).build()
    ps = lp_b200.presolve(pb, scale_passes=3)
    assert ps.dropped_duplicate == 2 and ps.dropped_empty == 1 and ps.problem.A().shape == (m0, len(c))
    A2 = np.abs(ps.problem.A())
    spread = lambda M: np.log10(M[M > 0].max() / M[M > 0].min())
    assert spread(A2) < spread(np.abs(pb.A())) - 3           # rows / columns pulled together by > 3 orders
    # the oracle on the presolved problem, mapped back, solves the original LP
    ref_y = o.InteriorPoint().solve(o.Problem(ps.problem.A(), ps.problem.b(), ps.problem.c(), 0.0, 0))
    back = ps.restore(lp_b200.OptimizeResult(ref_y.x, ref_y.fun, ref_y.iteration))
    np.testing.assert_allclose(A_eq @ back.x(), b_eq, rtol=1e-7, atol=1e-7 * np.abs(b_eq).max())
    assert (back.x() >= -1e-9).all()
    assert abs(c @ back.x() - back.fun()) <= 1e-7 * max(1.0, abs(back.fun()))
    with pytest.raises(o.LinearProgramError):                # the unpresolved problem: M is exactly singular
        o.InteriorPoint().solve(o.Problem(pb.A(), pb.b(), pb.c(), 0.0, 0))


def test_presolve_detects_contradicting_rows():
    c, A_eq, b_eq, _ = _messy_lp(1)
    b_bad = b_eq.copy()
    b_bad[-3] += 1.0                                         # the duplicate of row 3 now disagrees with it
    with pytest.raises(lp_b200.Infeasible):
        lp_b200.presolve(lp_b200.Problem.target(c).eq(A_eq, b_bad).build())
    b_bad = b_eq.copy()
    b_bad[-1] = 2.0                                          # 0 . x = 2
    with pytest.raises(lp_b200.Infeasible):
        lp_b200.presolve(lp_b200.Problem.target(c).eq(A_eq, b_bad).build())


def test_presolve_leaves_a_clean_problem_alone():
    args = o.synthetic_lp(64, 128, 0)
    pb = lp_b200.Problem.target(args[0]).ub(args[1], args[2]).eq(args[3], args[4]).build()
    ps = lp_b200.presolve(pb, scale_passes=0)
    assert ps.dropped_duplicate == 0 and ps.dropped_empty == 0
    np.testing.assert_array_equal(ps.problem.A(), pb.A())
    np.testing.assert_array_equal(ps.problem.b(), pb.b())
    assert ps.problem.n_slack() == pb.n_slack()


@pytest.mark.gpu
def test_presolved_problem_solves_on_the_gpu():
    c, A_eq, b_eq, _ = _messy_lp(2)
    ps = lp_b200.presolve(lp_b200.Problem.target(c).eq(A_eq, b_eq).build(), scale_passes=3)
    res = ps.restore(lp_b200.InteriorPoint.default().solve(ps.problem))
    ref_y = o.InteriorPoint().solve(o.Problem(ps.problem.A(), ps.problem.b(), ps.problem.c(), 0.0, 0))
    ref = ps.restore(lp_b200.OptimizeResult(ref_y.x, ref_y.fun, ref_y.iteration))
    assert abs(res.iteration() - ref.iteration()) <= 1
    assert abs(res.fun() - ref.fun()) <= 1e-8 * max(1.0, abs(ref.fun()))
    np.testing.assert_allclose(A_eq @ res.x(), b_eq, rtol=1e-6, atol=1e-6 * np.abs(b_eq).max())
