"""K6 batched mode (one CTA per LP) on the B200 vs the oracle looped over the batch, and vs the
single-problem GPU path.  Same parity bar per problem: status equal, iterations +-1, x 1e-6, fun 1e-8 rel."""
import numpy as np
import pytest

import lp_b200
from lp_b200 import _ffi
from oracle import ipm_oracle as o

pytestmark = pytest.mark.gpu


def make_batch(batch, m, n, seed0):
    As, bs, cs = [], [], []
    for i in range(batch):
        pb = o.build_problem(*o.synthetic_lp(m, n, seed0 + i))
        As.append(pb.A)
        bs.append(pb.b)
        cs.append(pb.c)
    return np.stack(As), np.stack(bs), np.stack(cs), m // 2


@pytest.mark.parametrize("batch,m,n", [(64, 64, 128), (7, 10, 30), (33, 64, 100), (5, 2, 4)])
def test_batched_matches_oracle(batch, m, n):
    A, b, c, n_slack = make_batch(batch, m, n, 1000)  # C4 uses seeds 1000+i
    res = lp_b200.solve_batched(A, b, c, n_slack=n_slack)
    for i in range(batch):
        ref = o.InteriorPoint().solve(o.Problem(A[i], b[i], c[i], 0.0, n_slack))
        assert res.status[i] == _ffi.LPB_OK
        assert abs(int(res.iteration[i]) - ref.iteration) <= 1
        assert np.abs(res.x[i] - ref.x).max() < 1e-6
        assert abs(res.fun[i] - ref.fun) <= 1e-8 * max(1.0, abs(ref.fun))


def test_batched_equals_single_problem_path():
    A, b, c, n_slack = make_batch(8, 64, 128, 2000)
    res = lp_b200.solve_batched(A, b, c, n_slack=n_slack)
    for i in range(8):
        m2 = 32
        pb = (lp_b200.Problem.target(c[i][: 128 - n_slack]).ub(A[i][:m2, : 128 - n_slack], b[i][:m2])
              .eq(A[i][m2:, : 128 - n_slack], b[i][m2:]).build())
        single = lp_b200.InteriorPoint.default().solve(pb)
        assert abs(int(res.iteration[i]) - single.iteration()) <= 1
        np.testing.assert_allclose(res.x[i], single.x(), rtol=0, atol=1e-7)
        assert abs(res.fun[i] - single.fun()) <= 1e-8 * max(1.0, abs(single.fun()))


def test_batched_slack_structure_is_detected_from_the_data_not_assumed():
    """The batched kernel does not store the trailing identity block of slack-form problems (then two CTAs fit on an
    SM).  The structure is read off the data per batch: (a) the same LPs with their columns permuted -- slack columns
    first, so there is no trailing identity -- take the store-everything path and must give the same solutions;
    (b) a batch in which ONE problem breaks the pattern (a 2 in the identity block) falls back as a whole.  (Batches
    with other numbers of inequality rows, i.e. other run lengths, are the parametrised cases above.)"""
    A, b, c, n_slack = make_batch(6, 64, 128, 4000)
    ref = lp_b200.solve_batched(A, b, c, n_slack=n_slack)
    assert all(ref.status == _ffi.LPB_OK)
    perm = np.concatenate([np.arange(128 - n_slack, 128), np.arange(128 - n_slack)])
    got = lp_b200.solve_batched(A[:, :, perm], b, c[:, perm])
    assert np.array_equal(got.iteration, ref.iteration) or np.abs(got.iteration - ref.iteration).max() <= 1
    np.testing.assert_allclose(got.x_slack[:, np.argsort(perm)], ref.x_slack, rtol=0, atol=1e-7)
    np.testing.assert_allclose(got.fun, ref.fun, rtol=1e-9)
    A2 = A.copy()
    A2[3, 5, 128 - n_slack + 5] = 2.0             # problem 3: that slack column is 2 e_5 now, a different (still valid) LP
    mixed = lp_b200.solve_batched(A2, b, c, n_slack=n_slack)
    one = o.InteriorPoint().solve(o.Problem(A2[3], b[3], c[3], 0.0, n_slack))
    assert abs(int(mixed.iteration[3]) - one.iteration) <= 1 and np.abs(mixed.x[3] - one.x).max() < 1e-6
    keep = [i for i in range(6) if i != 3]
    np.testing.assert_allclose(mixed.x_slack[keep], ref.x_slack[keep], rtol=0, atol=1e-7)


def test_batched_statuses_and_limits():
    # problem 0 infeasible (x1 + x2 + s = -1), problem 1 unbounded, problem 2 fine; all 1 x 3 slack form
    A = np.array([[[1.0, 1.0, 1.0]], [[1.0, -1.0, 1.0]], [[1.0, 1.0, 1.0]]])
    b = np.array([[-1.0], [1.0], [2.0]])
    c = np.array([[1.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [-1.0, -2.0, 0.0]])
    res = lp_b200.solve_batched(A, b, c, n_slack=1)
    assert list(res.status) == [_ffi.LPB_ERR_INFEASIBLE, _ffi.LPB_ERR_UNBOUNDED, _ffi.LPB_OK]
    np.testing.assert_allclose(res.x[2], [0.0, 2.0], atol=1e-6)
    res = lp_b200.solve_batched(A, b, c, n_slack=1, solver=lp_b200.InteriorPoint.custom().max_iter(1).build())
    assert res.status[2] == _ffi.LPB_ERR_ITERATION_LIMIT_EXCEEDED and res.iteration[2] == 1
    with pytest.raises(lp_b200.InvalidParameter):
        lp_b200.solve_batched(np.zeros((1, 65, 70)), np.zeros((1, 65)), np.zeros((1, 70)))  # m > 64 unsupported


def test_batched_staging_arena_and_pinned_inputs():
    """Host inputs are staged through a cached device arena: repeated calls, a call after
    lpb_release_workspaces and a call with page-locked inputs (lp_b200.pinned_empty) all give the same bits."""
    A, b, c, n_slack = make_batch(16, 64, 128, 3000)
    r1 = lp_b200.solve_batched(A, b, c, n_slack=n_slack)
    r2 = lp_b200.solve_batched(A, b, c, n_slack=n_slack)
    assert _ffi.load().lpb_release_workspaces() == 0
    r3 = lp_b200.solve_batched(A[:5], b[:5], c[:5], n_slack=n_slack)       # smaller batch: arena re-created
    Ap, bp, cp = lp_b200.pinned_empty(A.shape), lp_b200.pinned_empty(b.shape), lp_b200.pinned_empty(c.shape)
    Ap[:], bp[:], cp[:] = A, b, c
    r4 = lp_b200.solve_batched(Ap, bp, cp, n_slack=n_slack)
    for r in (r2, r4):
        np.testing.assert_array_equal(r.x, r1.x)
        np.testing.assert_array_equal(r.iteration, r1.iteration)
    np.testing.assert_array_equal(r3.x, r1.x[:5])
    assert (r1.status == _ffi.LPB_OK).all()
