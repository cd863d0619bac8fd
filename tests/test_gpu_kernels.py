"""Kernel-level parity on the B200, through the C ABI (lpb_k_*), against NumPy/LAPACK on the host.

FP64 throughout; tolerances are relative to the magnitude of the exact result and stated per test.
"""
import ctypes as C

import numpy as np
import pytest

from lp_b200 import _ffi
from tests.gpu_util import BareCtx, ok, pad_cols, to_dev

pytestmark = pytest.mark.gpu

# the last three cross K1's blocked-accumulation boundaries (a row group of the slab is folded into C every 2048 columns,
# first at column 256 (mi + 1)): one flush per group, a ragged tail behind a flush, three flushes per group
# ... and the last two have more tiles than SMs with a short last wave (153 tiles: 5 left over; 171: 23), so the tail tiles
# are cut into K-parts with private slots and summed by syrk_tail_fixup_kernel (dmma_gemm.cu: TailSplit)
SYRK_SHAPES = [(8, 16), (100, 250), (128, 256), (129, 257), (300, 1000), (513, 1031), (1024, 2048), (257, 2320),
               (129, 4130), (200, 6500), (2176, 1024), (2200, 1100)]


@pytest.mark.parametrize("impl", [0, 1])
@pytest.mark.parametrize("m,n", SYRK_SHAPES)
@pytest.mark.parametrize("scaled", [True, False])
def test_syrk_adat_matches_numpy(m, n, impl, scaled):
    """K1 vs newton_equations.rs:54-57 computed with NumPy; rel. tol 1e-12 of |A| D |A|^T."""
    rng = np.random.default_rng(m * 1000 + n)
    A = rng.standard_normal((m, n))
    d = np.exp(rng.uniform(-6, 6, n)) if scaled else None
    Ap, lda = pad_cols(A)
    ldm = (m + 15) // 16 * 16
    dA = to_dev(Ap)
    dd = to_dev(d) if scaled else None
    import torch
    dM = torch.full((m, ldm), float("nan"), dtype=torch.float64, device="cuda")
    with BareCtx(m, n) as ctx:
        ctx.set("syrk_impl", impl)
        ok(ctx.lib.lpb_k_syrk_adat(ctx.h, m, n, dA.data_ptr(), lda, dd.data_ptr() if scaled else None,
                                   dM.data_ptr(), ldm))
    M = dM.cpu().numpy()[:, :m]
    D = d if scaled else np.ones(n)
    ref = (A * D) @ A.T
    scale = (np.abs(A) * D) @ np.abs(A).T
    low = np.tril_indices(m)
    err = np.abs(M[low] - ref[low]) / scale[low]
    assert np.isfinite(M[low]).all()
    assert err.max() < 1e-12


@pytest.mark.parametrize("period", [32, 64, 128, 1024])
@pytest.mark.parametrize("m,n", [(129, 4130), (300, 2049), (64, 513)])
def test_syrk_every_flush_period_gives_the_same_matrix_up_to_rounding(m, n, period):
    """Option "syrk_flush_blocks": the period of K1's blocked accumulation (K-blocks of 16 columns; a power of two
    >= 32; longer than the K extent = never).  Shapes put the first / second flush of every row group, the ragged last
    K-block and the half-period offset of the second four warps on and next to a boundary."""
    import torch
    rng = np.random.default_rng(m + n + period)
    A = rng.standard_normal((m, n))
    d = np.exp(rng.uniform(-3, 3, n))
    Ap, lda = pad_cols(A)
    dA, dd = to_dev(Ap), to_dev(d)
    dM = torch.full((m, m + (m & 1)), float("nan"), dtype=torch.float64, device="cuda")
    with BareCtx(m, n) as ctx:
        ctx.set("syrk_flush_blocks", period)
        assert ctx.lib.lpb_set_option(ctx.h, b"syrk_flush_blocks", 48) != 0       # not a power of two
        assert ctx.lib.lpb_set_option(ctx.h, b"syrk_flush_blocks", 16) != 0       # too short
        ok(ctx.lib.lpb_k_syrk_adat(ctx.h, m, n, dA.data_ptr(), lda, dd.data_ptr(), dM.data_ptr(), m + (m & 1)))
    M = dM.cpu().numpy()[:, :m]
    ref = (A * d) @ A.T
    scale = (np.abs(A) * d) @ np.abs(A).T
    low = np.tril_indices(m)
    assert np.isfinite(M[low]).all()
    assert (np.abs(M[low] - ref[low]) / scale[low]).max() < 1e-13


def test_syrk_tail_split_changes_nothing_but_the_rounding_of_the_tail_tiles():
    """Option "syrk_tail_split" = 0 (every work item a whole tile) against the default on a shape whose last wave holds 23
    of 171 tiles: the same matrix up to rounding, every entry finite, and bit-identical from run to run (the parts are
    summed in a fixed order: no atomics)."""
    import torch
    m, n = 2200, 1100
    rng = np.random.default_rng(5)
    A = rng.standard_normal((m, n))
    d = np.exp(rng.uniform(-3, 3, n))
    Ap, lda = pad_cols(A)
    dA, dd = to_dev(Ap), to_dev(d)
    out = []
    for split in (1, 1, 0):
        dM = torch.full((m, m), float("nan"), dtype=torch.float64, device="cuda")
        with BareCtx(m, n) as ctx:
            ctx.set("syrk_tail_split", split)
            ok(ctx.lib.lpb_k_syrk_adat(ctx.h, m, n, dA.data_ptr(), lda, dd.data_ptr(), dM.data_ptr(), m))
        out.append(np.tril(dM.cpu().numpy()))
    assert np.isfinite(out[0]).all()
    assert np.array_equal(out[0], out[1])
    scale = np.tril((np.abs(A) * d) @ np.abs(A).T) + 1e-300
    assert (np.abs(out[0] - out[2]) / scale).max() < 1e-13
    assert not np.array_equal(out[0], out[2])     # the tail tiles really took the other path


def test_syrk_blocked_accumulation_keeps_long_same_sign_sums_to_a_few_ulp():
    """The diagonal of M = A D A^T is a sum of n same-sign terms.  One register chain of n / 4 DMMA steps rounds
    relative to the growing partial sum (measured 26 ulp rms at n = 24576: option "syrk_chain" = 1, the round-1
    kernel, and cuBLAS DGEMM behave alike); K1 folds its accumulators into C every 2048 columns, which a BLAS that
    blocks K does implicitly (the CPU oracle's OpenBLAS).  Checked against the diagonal summed in extended precision."""
    import torch
    m, n = 256, 24576
    rng = np.random.default_rng(7)
    A = rng.standard_normal((m, n))
    d = np.exp(rng.uniform(-1, 1, n))
    exact = ((A.astype(np.longdouble) ** 2) * d.astype(np.longdouble)).sum(axis=1)
    dA, dd = to_dev(A), to_dev(d)
    rms = {}
    for chain in (0, 1):
        dM = torch.full((m, m), float("nan"), dtype=torch.float64, device="cuda")
        with BareCtx(m, n) as ctx:
            ctx.set("syrk_chain", chain)
            ok(ctx.lib.lpb_k_syrk_adat(ctx.h, m, n, dA.data_ptr(), n, dd.data_ptr(), dM.data_ptr(), m))
        got = np.diagonal(dM.cpu().numpy()).astype(np.longdouble)
        ulps = np.abs(got - exact) / np.spacing(exact.astype(np.float64))
        rms[chain] = float(np.sqrt((ulps.astype(np.float64) ** 2).mean()))
    print("diagonal of M, n = %d same-sign terms: %.2f ulp rms blocked, %.2f ulp rms as one chain" % (n, rms[0], rms[1]))
    assert rms[0] < 4.0 and rms[0] < 0.5 * rms[1]


@pytest.mark.parametrize("impl,trsm", [(0, 0), (0, 1), (0, 2), (0, 3), (1, 0)])
@pytest.mark.parametrize("m", [1, 5, 16, 17, 64, 128, 129, 200, 384, 1000, 1536])
def test_potrf_matches_lapack(m, impl, trsm):
    """K2 vs numpy.linalg.cholesky (LAPACK potrf); ||L L^T - M|| / ||M|| < 1e-13 and L close to LAPACK's."""
    rng = np.random.default_rng(m)
    B = rng.standard_normal((m, m + 8))
    M = B @ B.T + 0.1 * np.eye(m)
    Mp, ldm = pad_cols(M)
    dM = to_dev(Mp)
    info = C.c_int32(-1)
    with BareCtx(m, m) as ctx:
        ctx.set("syrk_impl", impl)
        ctx.set("trsm_impl", trsm)  # 0 blocked substitution on DMMA (default), 1 column substitution, 2 GEMM with inv(L_kk), 3 blocked substitution in DFMA
        ok(ctx.lib.lpb_k_potrf(ctx.h, m, dM.data_ptr(), ldm, C.byref(info)))
    assert info.value == 0
    L = np.tril(dM.cpu().numpy()[:, :m])
    assert np.linalg.norm(L @ L.T - M) / np.linalg.norm(M) < 1e-13
    Lref = np.linalg.cholesky(M)
    assert np.abs(L - Lref).max() / np.abs(Lref).max() < 1e-10


def test_potrf_reports_non_positive_pivot():
    """newton_equations.rs:63: a failed factorisation must surface (info = first bad pivot + 1)."""
    m = 300
    rng = np.random.default_rng(7)
    B = rng.standard_normal((m, m))
    M = B @ B.T + np.eye(m)
    M[200, 200] = -1.0
    Mp, ldm = pad_cols(M)
    dM = to_dev(Mp)
    info = C.c_int32(0)
    with BareCtx(m, m) as ctx:
        ok(ctx.lib.lpb_k_potrf(ctx.h, m, dM.data_ptr(), ldm, C.byref(info)))
    assert info.value == 201


@pytest.mark.parametrize("nrhs", [1, 2])
@pytest.mark.parametrize("m", [7, 128, 130, 500, 1536])
def test_potrs_matches_lapack(m, nrhs):
    """K3 vs scipy cho_solve; rel. error < 1e-10 on a well conditioned system."""
    from scipy.linalg import cho_solve
    rng = np.random.default_rng(m + nrhs)
    Bm = rng.standard_normal((m, m + 8))
    M = Bm @ Bm.T + m * np.eye(m)
    L = np.linalg.cholesky(M)
    Lp, ldm = pad_cols(L)
    rhs = rng.standard_normal((nrhs, m))  # column-major m x nrhs == row-major nrhs x m
    dL, dB = to_dev(Lp), to_dev(rhs)
    with BareCtx(m, m) as ctx:
        ok(ctx.lib.lpb_k_potrs(ctx.h, m, dL.data_ptr(), ldm, dB.data_ptr(), nrhs))
    X = dB.cpu().numpy()
    ref = cho_solve((L, True), rhs.T).T
    assert np.abs(X - ref).max() / np.abs(ref).max() < 1e-10


@pytest.mark.parametrize("m,n", [(3, 5), (64, 128), (100, 251), (512, 1024), (1000, 4097)])
def test_gemv_sweeps_match_numpy(m, n):
    """K4: A w and A^T v, rel. tol 1e-13 of |A||w|."""
    rng = np.random.default_rng(n)
    A = rng.standard_normal((m, n))
    w = rng.standard_normal(n)
    v = rng.standard_normal(m)
    Ap, lda = pad_cols(A)
    dA, dw, dv = to_dev(Ap), to_dev(w), to_dev(v)
    import torch
    o_n = torch.zeros(m, dtype=torch.float64, device="cuda")
    o_t = torch.zeros(n, dtype=torch.float64, device="cuda")
    with BareCtx(m, n) as ctx:
        ok(ctx.lib.lpb_k_gemv_n(ctx.h, m, n, dA.data_ptr(), lda, dw.data_ptr(), o_n.data_ptr()))
        ok(ctx.lib.lpb_k_gemv_t(ctx.h, m, n, dA.data_ptr(), lda, dv.data_ptr(), o_t.data_ptr()))
    assert (np.abs(o_n.cpu().numpy() - A @ w) / (np.abs(A) @ np.abs(w))).max() < 1e-13
    assert (np.abs(o_t.cpu().numpy() - A.T @ v) / (np.abs(A.T) @ np.abs(v))).max() < 1e-13


@pytest.mark.parametrize("solve_impl,grid_cap", [(0, 0), (0, 3), (0, 1), (1, 0), (2, 0), (3, 0), (3, 3), (3, 1)])
@pytest.mark.parametrize("nrhs", [1, 2])
@pytest.mark.parametrize("m", [1, 17, 64, 128, 200, 640, 1537, 2305])
def test_potrf_then_fused_potrs(m, nrhs, solve_impl, grid_cap):
    """K2 + K3 fast path: factor on the device, then solve with the stored inverted diagonal blocks
    (solve_impl 0: the single-launch kernel with tagged hand-off and full block inverses, also with its grid capped
    so one CTA owns several block rows; 1: one launch per block; 2: plain substitution; 3: the flag-based pipelined
    kernel with blocked substitution); residual ||M x - b|| / (||M|| ||x||) < 1e-13, x close to LAPACK's."""
    from scipy.linalg import cho_factor, cho_solve
    rng = np.random.default_rng(3 * m + nrhs)
    Bm = rng.standard_normal((m, m + 8))
    M = Bm @ Bm.T + 0.5 * np.eye(m)
    Mp, ldm = pad_cols(M)
    rhs = rng.standard_normal((nrhs, m))
    dM, dB = to_dev(Mp), to_dev(rhs)
    info = C.c_int32(-1)
    with BareCtx(m, m) as ctx:
        ctx.set("solve_impl", solve_impl)
        ctx.set("solve_grid_cap", grid_cap)
        ok(ctx.lib.lpb_k_potrf(ctx.h, m, dM.data_ptr(), ldm, C.byref(info)))
        assert info.value == 0
        ok(ctx.lib.lpb_k_potrs(ctx.h, m, dM.data_ptr(), ldm, dB.data_ptr(), nrhs))
        if solve_impl in (0, 3):  # a second solve on the same context: the hand-off epoch advances
            dB2 = to_dev(rhs)
            ok(ctx.lib.lpb_k_potrs(ctx.h, m, dM.data_ptr(), ldm, dB2.data_ptr(), nrhs))
            assert np.array_equal(dB2.cpu().numpy(), dB.cpu().numpy())
    X = dB.cpu().numpy()
    ref = cho_solve(cho_factor(M, lower=True), rhs.T).T
    assert np.abs(X - ref).max() / np.abs(ref).max() < 1e-9
    for k in range(nrhs):
        assert np.linalg.norm(M @ X[k] - rhs[k]) / (np.linalg.norm(M) * np.linalg.norm(X[k])) < 1e-13


@pytest.mark.parametrize("m", [257, 384, 1000, 1537, 4096])
def test_potrf_lookahead_is_bit_identical_to_the_sequential_loop(m):
    """K2 with look-ahead (potf2 of panel k+1 on a second stream beside the trailing update of panel k) applies
    the same updates to every tile in the same order as the sequential loop: identical bits, twice in a row."""
    rng = np.random.default_rng(5 * m)
    B = rng.standard_normal((m, m + 8))
    M = B @ B.T + 0.1 * np.eye(m)
    Mp, ldm = pad_cols(M)
    outs = []
    info = C.c_int32(-1)
    with BareCtx(m, m) as ctx:
        for look in (1, 0, 1):
            ctx.set("potrf_lookahead", look)
            dM = to_dev(Mp)
            ok(ctx.lib.lpb_k_potrf(ctx.h, m, dM.data_ptr(), ldm, C.byref(info)))
            assert info.value == 0
            outs.append(np.tril(dM.cpu().numpy()[:, :m]))
    np.testing.assert_array_equal(outs[0], outs[1])
    np.testing.assert_array_equal(outs[0], outs[2])
    assert np.linalg.norm(outs[0] @ outs[0].T - M) / np.linalg.norm(M) < 1e-13
