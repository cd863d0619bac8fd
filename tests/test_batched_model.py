"""CPU models of two pieces of index algebra in lp_b200/csrc/batched.cu (K6), restated lane by lane in NumPy so that
a slip in them is caught without a GPU (the GPU parity tests are tests/test_gpu_batched.py):

* rows_dot: a warp sums its 8 rows together and folds the 8 x 32 partial sums with a transposing butterfly -- at
  offsets 16 / 8 / 4 each lane keeps half of its rows and adds the partner's partials of those, then two plain
  exchanges; lane l must end up with row (l >> 2) summed over all 32 lanes;
* the in-CTA SYRK on mma.m8n8k4 fragments over the lower-triangle 8 x 8 tiles, with the implicit slack block added on
  the diagonal by whichever lane holds that entry, and a zero-padded K tail.
"""
import numpy as np
import pytest


def butterfly_fold(part):
    """part[lane, r]: lane's partial sum of local row r (8 rows).  Returns out[lane] and the local row it belongs to."""
    s = part.copy()
    R = s.shape[1]
    half, off = R // 2, 16
    lanes = np.arange(32)
    while half >= 1:
        up = (lanes & off) != 0
        new = s.copy()
        for r in range(half):
            send = np.where(up, s[:, r], s[:, r + half])
            keep = np.where(up, s[:, r + half], s[:, r])
            new[:, r] = keep + send[lanes ^ off]          # __shfl_xor_sync(send, off)
        s = new
        half //= 2
        off //= 2
    v = s[:, 0]
    v = v + v[lanes ^ 2]
    v = v + v[lanes ^ 1]
    return v, lanes >> 2


def test_transposing_butterfly_gives_every_lane_the_full_sum_of_row_lane_over_4():
    rng = np.random.default_rng(0)
    part = rng.standard_normal((32, 8))
    out, row = butterfly_fold(part)
    want = part.sum(axis=0)
    for lane in range(32):
        assert abs(out[lane] - want[row[lane]]) <= 1e-13 * np.abs(part).sum()
    # exactly the same bits on the four lanes that share a row (the kernel lets lane & 3 == 0 write)
    for lane in range(0, 32, 4):
        assert len({out[lane + e] for e in range(4)}) == 1


def dmma_tile(Ablk, Bblk, d):
    """One 8 x 8 output tile accumulated over K by m8n8k4 fragments: lane (g, t) feeds A[g][k0 + t] and
    B[k0 + t][g] = Bblk[g][k0 + t] * d[k0 + t]; it owns C[g][2 t], C[g][2 t + 1]."""
    K = Ablk.shape[1]
    C = np.zeros((8, 8))
    for k0 in range(0, K, 4):
        a = np.array([[Ablk[g, k0 + t] for t in range(4)] for g in range(8)])          # per-lane A values
        b = np.array([[Bblk[g, k0 + t] * d[k0 + t] for t in range(4)] for g in range(8)])
        C += a @ b.T                                                                    # what the 32 lanes' mma adds up to
    return C


@pytest.mark.parametrize("m,nd,ns", [(64, 96, 32), (10, 25, 5), (37, 50, 0), (8, 3, 8)])
def test_fragment_syrk_covers_the_lower_triangle_and_adds_the_slack_diagonal(m, nd, ns):
    rng = np.random.default_rng(m + nd)
    lda = (nd + 15) // 16 * 16 + 4                       # batched_lda
    A = np.zeros((((m + 7) // 8) * 8, lda))               # zero K padding; rows >= m may hold anything
    A[:m, :nd] = rng.standard_normal((m, nd))
    A[m:, :] = np.nan
    dinv = np.exp(rng.uniform(-2, 2, nd + ns))
    dpad = np.concatenate([dinv[:nd], np.full(lda - nd, dinv[nd - 1])])     # index clamped to nd - 1
    nt = (m + 7) // 8
    M = np.full((m, m), np.nan)
    for bi in range(nt):
        for bj in range(bi + 1):
            kpad = (nd + 3) // 4 * 4
            with np.errstate(invalid="ignore"):
                C = dmma_tile(A[bi * 8: bi * 8 + 8, :kpad], A[bj * 8: bj * 8 + 8, :kpad], dpad[:kpad])
            for g in range(8):
                for t in range(4):
                    row, col = bi * 8 + g, bj * 8 + 2 * t
                    if row >= m:
                        continue
                    c0, c1 = C[g, 2 * t], C[g, 2 * t + 1]
                    if row < ns and row == col:
                        c0 += dinv[nd + row]
                    if row < ns and row == col + 1:
                        c1 += dinv[nd + row]
                    if col < m:
                        M[row, col] = c0
                    if col + 1 < m:
                        M[row, col + 1] = c1
    full = np.hstack([A[:m, :nd], np.eye(m)[:, :ns]])
    want = (full * dinv) @ full.T
    low = np.tril_indices(m)
    assert np.isfinite(M[low]).all()
    assert np.abs(M[low] - want[low]).max() <= 1e-12 * np.abs(want).max()
