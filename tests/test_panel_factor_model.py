"""CPU model of lp_b200/csrc/panel_factor.cuh (the 16 x 16 pivot loop shared by K2's panel kernel and the batched
kernel): the arithmetic of the kernel restated operation by operation in Python, so that the two ideas it rests on
are checked without a GPU --

* pivot_rsqrt: a ~2^-22 seed (MUFU.RSQ64H reads only the high word of its argument) followed by ONE third-order step
  y (1 + e/2 + 3 e^2 / 8), e = 1 - a y^2, is within 1 ulp of 1 / sqrt(a) over the whole exponent range;
* factor_sub16: the ROLLED, SHIFTING loop -- at pivot j, v[k] holds the entry of column j + k, the update
  v[k] <- v[k+1] - l * L[j+1+k][j] moves the row down by one -- factors the block (lanes 0..15) and inverts it in the
  same 16 steps (lanes 16..31 run the forward substitutions L x = e_i), and a bad pivot poisons its lane so that one
  test per sub-block finds the first one.

The GPU parity tests (tests/test_gpu_kernels.py) check the kernel itself against LAPACK; this file checks the algorithm.
"""
import struct
from fractions import Fraction

import numpy as np
import pytest


def fma(a, b, c):
    """Correctly rounded a * b + c (exact rational arithmetic, one rounding)."""
    return float(Fraction(a) * Fraction(b) + Fraction(c))


def high_word_only(x):
    """The double whose low 32 bits are zero (what an instruction that reads the high word sees / returns)."""
    (u,) = struct.unpack("<Q", struct.pack("<d", x))
    return struct.unpack("<d", struct.pack("<Q", u & 0xFFFFFFFF00000000))[0]


def pivot_rsqrt(a):
    if not (a > 0.0 and np.isfinite(a)):          # the hardware seed returns NaN / inf here and the step keeps it bad
        return float("nan")
    y = high_word_only(1.0 / np.sqrt(high_word_only(a)))      # the seed: ~2^-21 relative
    e = fma(-(a * y), y, 1.0)
    return fma(y * e, fma(0.375, e, 0.5), y)


def test_seed_plus_one_cubic_step_is_within_one_ulp():
    rng = np.random.default_rng(0)
    worst = 0.0
    for _ in range(3000):
        a = float(np.ldexp(rng.uniform(1.0, 2.0), int(rng.integers(-600, 600))))
        y = pivot_rsqrt(a)
        exact = 1 / Fraction(a)                      # y^2 should be 1 / a: compare squares exactly, no sqrt of a Fraction
        # |y - a^-1/2| / ulp via the exact identity  y - r = (y^2 - r^2) / (y + r),  r ~ y
        err = abs(Fraction(y) ** 2 - exact) / (2 * Fraction(y))
        worst = max(worst, float(err / Fraction(np.spacing(y))))
    assert worst <= 1.0, worst


def factor_sub16(Ablk, rsqrt=pivot_rsqrt):
    """The warp's 32 lanes as 32 rows of a NumPy array; returns (L, X, bad_mask)."""
    SB = 16
    v = np.zeros((32, SB))
    dg = np.ones(32)
    for lane in range(32):
        i = lane & 15
        if lane < SB:
            v[lane, : i + 1] = Ablk[i, : i + 1]
            dg[lane] = Ablk[i, i]
        else:
            v[lane, i] = 1.0
    L = np.zeros((SB, SB))
    X = np.zeros((SB, SB))
    myrd = np.ones(32)
    with np.errstate(all="ignore"):
        rs = rsqrt(dg[0])
        for j in range(SB):
            l = v[:, 0] * rs
            myrd[j] = rs                                      # lane j (factor half) keeps 1 / L[j][j]
            dg = dg - l * l                                   # fma(-l, l, dg) on the GPU
            rs_next = rsqrt(dg[(j + 1) & 15])                 # shuffle from lane j + 1, then the seed + cubic step
            cb = np.concatenate([l[:SB], np.full(SB, np.nan)])  # column j of L; the padding half is never consumed
            for lane in range(32):
                i = lane & 15
                if lane < SB and j <= i:
                    L[i, j] = l[lane]
                if lane >= SB:
                    X[j, i] = l[lane] if j >= i else 0.0      # X[r = j][c = i]
            v[:, : SB - 1] = v[:, 1:] - l[:, None] * cb[j + 1: j + SB][None, :]   # the shift
            v[:, SB - 1] = np.nan
            rs = rs_next
    bad = [not (myrd[i] > 0.0 and np.isfinite(myrd[i])) for i in range(SB)]
    return L, X, bad


def _rsqrt_plain(a):
    with np.errstate(all="ignore"):
        return float(1.0 / np.sqrt(a)) if a == a else float("nan")


@pytest.mark.parametrize("seed", range(4))
def test_shifting_pivot_loop_factors_and_inverts(seed):
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((16, 40))
    A = B @ np.diag(np.exp(rng.uniform(-4, 4, 40))) @ B.T
    L, X, bad = factor_sub16(A, rsqrt=_rsqrt_plain if seed else pivot_rsqrt)
    assert not any(bad)
    assert np.allclose(L, np.linalg.cholesky(A), rtol=1e-10, atol=0)
    assert np.abs(L @ L.T - A).max() <= 1e-14 * np.abs(A).max()
    assert np.abs(X @ L - np.eye(16)).max() <= 1e-9
    assert np.array_equal(np.triu(X, 1), np.zeros((16, 16)))   # X is lower triangular, zeros stored above


def test_a_bad_pivot_poisons_its_lane_and_every_later_one():
    rng = np.random.default_rng(9)
    B = rng.standard_normal((16, 30))
    A = B @ B.T
    A[5, 5] = -abs(A[5, 5])      # the Schur complement turns negative at pivot 5 at the latest
    _, _, bad = factor_sub16(A, rsqrt=_rsqrt_plain)
    first = bad.index(True)
    assert first <= 5 and all(bad[first:])         # one test per sub-block finds the FIRST bad pivot: ffs of the mask
    assert not any(bad[:first])
