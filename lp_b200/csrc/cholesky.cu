// cholesky.cu -- K2: blocked right-looking lower Cholesky of M (row-major, in place) and
// K3: the forward/back substitutions, 1 or 2 right-hand sides at once.
//
// Replaces /root/reference/src/solvers/interior_point/newton_equations.rs:129-132 (`M.cholesky()`,
// lower L) / :87-90 (LAPACK potrf) and :151-169 (`factor.solvec_into`) / :98-104 (potrs).
// A non-positive or non-finite pivot sets the device `info` flag -> LPB_ERR_NUMERICAL_PROBLEM
// (newton_equations.rs:63); there is no fallback chain.
//
// Per 128-wide panel:  potf2_inv (one CTA: factor the diagonal block in shared memory and invert
// it in place, quad-per-column)  ->  panel TRSM as a DMMA GEMM  P <- P inv(L_kk)^T  ->  trailing
// update C -= P P^T on the DMMA path (dmma_gemm.cu).  The inverted diagonal blocks are kept
// (m x 128 doubles) so the solves need no serial substitution: every 128-block step is ONE launch
// in which each CTA forms x_k = inv(L_kk) b_k itself (a 128 x 128 GEMV out of L2) and applies the
// rank-128 update to its own rows / columns.
// The substitution-based trsm / trsv kernels below are the plain reference path (syrk_impl = 1).
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <mutex>
#include <type_traits>

#include "dist_schedule.hpp"
#include "kernels.hpp"
#include "panel_factor.cuh"

namespace lpb {

namespace {

constexpr int NB = 128;        // panel width == diagonal block
constexpr int LDS = NB + 1;    // padded smem pitch: column walks are bank-conflict free
constexpr int TRSM_ROWS = 64;  // rows of the panel per TRSM CTA

// ------------------------------------------------------------------ potf2: one CTA, nb <= 128
__global__ void __launch_bounds__(512)
potf2_kernel(double* __restrict__ Mat, int64_t ldm, int k0, int nb, int* __restrict__ info) {
  extern __shared__ double S[];  // NB*LDS block + NB diag
  double* diag = S + NB * LDS;
  const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
  double* blk = Mat + (int64_t)k0 * ldm + k0;
  for (int i = ty; i < nb; i += 16)
    for (int k = tx; k <= i; k += 32) S[i * LDS + k] = blk[(int64_t)i * ldm + k];
  __syncthreads();

  for (int j = 0; j < nb; ++j) {
    const double ajj = S[j * LDS + j];
    if (!(ajj > 0.0) || !isfinite(ajj)) {  // uniform: every thread reads the same shared value
      if (tid == 0 && *info == 0) *info = k0 + j + 1;
      // poison the rest of the block so the failure is visible downstream
      for (int i = j + tid; i < nb; i += 512) diag[i] = __longlong_as_double(0x7ff8000000000000ll);
      break;
    }
    const double d = sqrt(ajj);
    if (tid == 0) diag[j] = d;
    // scale column j (S[j][j] itself is left untouched: others are still reading it)
    if (tid < nb - j - 1) S[(j + 1 + tid) * LDS + j] /= d;
    __syncthreads();
    // rank-1 update of the trailing lower triangle
    for (int i = j + 1 + ty; i < nb; i += 16) {
      const double lij = S[i * LDS + j];
      for (int k = j + 1 + tx; k <= i; k += 32) S[i * LDS + k] -= lij * S[k * LDS + j];
    }
    __syncthreads();
  }
  __syncthreads();
  for (int i = ty; i < nb; i += 16)
    for (int k = tx; k <= i; k += 32) blk[(int64_t)i * ldm + k] = (k == i) ? diag[i] : S[i * LDS + k];
}

// ------------------------------------------------------------------ potf2 + inverse of the diagonal block (one CTA)
// This kernel sits on the critical path of every panel, so it is organised to keep the serial part
// short: the 128 x 128 block is factored in 16-column steps; per step ONE warp factors the 16 x 16
// diagonal sub-block entirely in registers (lane i = row i, pivots broadcast by shuffle) and inverts
// it, then all 16 warps apply X = inv(L_dd) to the rows below (a 16-wide GEMM, no substitution chain)
// and the rank-16 update to the remaining sub-blocks.  The full inverse X = inv(L_kk) is then built
// block-diagonal by block-diagonal:  X_ij = -X_ii (sum_{k=j}^{i-1} L_ik X_kj),  all blocks of one
// diagonal in parallel.  X^T lives in the strictly upper triangle of the same shared block (row c
// holds column c of X), its diagonal in rdiag.  Rows >= nb of a ragged last panel are padded with the
// identity so every loop is uniform.  L goes back to Mat, X to Linv (dense 128 x 128, zero above the
// diagonal and beyond nb).  A non-positive / non-finite pivot sets *info (first one wins) and poisons
// the block with NaN; there is no early exit.
constexpr int SB = 16;        // sub-block
constexpr int NSB = NB / SB;  // 8
constexpr int XDP = SB + 1;   // pitch of the 16 x 16 inverse of the current diagonal sub-block (odd: the four row groups a warp reads hit different banks)
constexpr int TWP = 9;        // pitch of the per-warp 16 x 8 scratch of the block-inverse step

// Off-diagonal 16 x 16 blocks of X = inv(L_kk) from L_kk (lower triangle of S) and the inverted diagonal sub-blocks
// (X^T in the strictly upper triangle of S, its diagonal in rdiag): one block diagonal at a time,
//     X_ij = -X_ii (sum_{k=j}^{i-1} L_ik X_kj),
// block (i, j) by a pair of warps; 512 threads, ends with a CTA barrier.
__device__ __forceinline__ void complete_block_inverse(double* S, const double* rdiag, double* Tw, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const int b = warp >> 1, h = warp & 1;
  const int ii = lane & 15, cg = lane >> 4;
  double* tw = Tw + warp * SB * TWP;
  for (int d = 1; d < NSB; ++d) {
    if (b < NSB - d) {
      const int j = b, i = b + d;
      const int ccb = 8 * h + 4 * cg;  // first of this lane's 4 columns within block column j
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      const double* lrow = S + (i * SB + ii) * LDS;
      // k == j: X_jj (lower triangular; transposed in the upper triangle, diagonal in rdiag)
#pragma unroll
      for (int l = 0; l < SB; ++l) {
        const double lv = lrow[j * SB + l];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int cc = ccb + e;
          const double up = S[(j * SB + cc) * LDS + j * SB + l];
          const double xv = (l > cc) ? up : ((l == cc) ? rdiag[j * SB + cc] : 0.0);
          acc[e] += lv * xv;
        }
      }
      for (int k = j + 1; k < i; ++k) {
#pragma unroll
        for (int l = 0; l < SB; ++l) {
          const double lv = lrow[k * SB + l];
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[e] += lv * S[(j * SB + ccb + e) * LDS + k * SB + l];
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) tw[ii * TWP + 4 * cg + e] = acc[e];
      __syncwarp();
      // X_ij = -X_ii T
      double o[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int l = 0; l < SB; ++l) {
        const double up = S[(i * SB + l) * LDS + i * SB + ii];
        const double xv = (l < ii) ? up : ((l == ii) ? rdiag[i * SB + ii] : 0.0);
#pragma unroll
        for (int e = 0; e < 4; ++e) o[e] += xv * tw[l * TWP + 4 * cg + e];
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) S[(j * SB + ccb + e) * LDS + i * SB + ii] = -o[e];
    }
    __syncthreads();
  }
}

#ifdef LPB_POTF2_PROF  // build-time aid: thread 0 (in the serial warp) time-stamps the phases of ONE launch and prints them
#define P2_T(k) do { if (tid == 0) p2t[k] = clock64(); } while (0)
#else
#define P2_T(k) do { } while (0)
#endif
__global__ void __launch_bounds__(512)
potf2_inv_kernel(double* __restrict__ Mat, int64_t ldm, int k0, int nb, int* __restrict__ info,
                 double* __restrict__ Linv, int full_inverse, double* __restrict__ pack_lkk = nullptr,
                 double* __restrict__ pack_linv = nullptr) {
  extern __shared__ double S[];    // NB*LDS block
  double* rdiag = S + NB * LDS;    // NB: 1 / L[i][i]
  double* Xd = rdiag + NB;         // 2 x SB * XDP: inv(L_dd) of the current and of the next diagonal sub-block
  double* Tw = Xd + 2 * SB * XDP;  // 16 warps * SB * TWP
  double* cbuf = Tw + 16 * SB * TWP;  // 2 x (SB + SB padding): column j of the diagonal sub-block, double-buffered; + SB * XDP scratch
  const unsigned full = 0xffffffffu;
  const int tid = threadIdx.y * 32 + threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(full, tid >> 5, 0);  // through a shuffle: the compiler then knows it is warp-uniform
#ifdef LPB_POTF2_PROF
  long long p2t[40];
#endif
  P2_T(0);
  double* blk = Mat + (int64_t)k0 * ldm + k0;
  // Warp 0 fetches the first 16 x 16 sub-block by itself and starts factoring it (after the definitions below) while
  // the other 15 warps bring in the rest: 16-byte loads (Mat is 16-byte aligned with an even pitch, k0 a multiple of
  // 128), identity padding beyond a ragged block.
  auto load_pair = [&](int r, int c) {  // entries (r, c), (r, c + 1), c even
    double2 v = make_double2(0.0, 0.0);
    const bool in = r < nb;
    if (in && c <= r) v = *reinterpret_cast<const double2*>(blk + (int64_t)r * ldm + c);
    S[r * LDS + c] = (in && c <= r) ? v.x : (r == c ? 1.0 : 0.0);
    S[r * LDS + c + 1] = (in && c + 1 <= r) ? v.y : (r == c + 1 ? 1.0 : 0.0);
  };
  if (warp == 0) {
#pragma unroll
    for (int it = 0; it < 4; ++it) load_pair(lane & 15, 2 * (lane >> 4) + 4 * it);
    __syncwarp();
  } else {
    // every load of a thread in flight at once: one L2 round trip instead of five
    constexpr int kIt = (NB * (NB / 2) + 479) / 480;
    double2 ld[kIt];
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int idx = tid - 32 + 480 * it, r = idx >> 6, c = 2 * (idx & 63);
      ld[it] = make_double2(0.0, 0.0);
      if (idx < NB * (NB / 2) && r < nb && c <= r && (r >= SB || c >= SB))
        ld[it] = *reinterpret_cast<const double2*>(blk + (int64_t)r * ldm + c);
    }
#pragma unroll
    for (int it = 0; it < kIt; ++it) {
      const int idx = tid - 32 + 480 * it, r = idx >> 6, c = 2 * (idx & 63);
      if (idx < NB * (NB / 2) && (r >= SB || c >= SB)) {
        const bool in = r < nb;
        S[r * LDS + c] = (in && c <= r) ? ld[it].x : (r == c ? 1.0 : 0.0);
        S[r * LDS + c + 1] = (in && c + 1 <= r) ? ld[it].y : (r == c + 1 ? 1.0 : 0.0);
      }
    }
  }

  // ---- the serial heart of the factorisation: ONE warp factors the 16 x 16 diagonal sub-block at c0 and inverts
  // it in the same 16 pivot steps.  Lanes 0..15 hold row i of the sub-block (v[k] = a[i][k]); lanes 16..31
  // hold the running sums of the forward substitution L x = e_i for column i of X = inv(L_dd)
  // (v[r] = sum_{l<r} L[r][l] x_l, replaced by x_r at step r).  Both need exactly column j of L at pivot
  // step j -- broadcast through shared memory -- and then the same FMA  v[k] += mult * L[k][j], k > j.
  // sqrt and the divisions are one rsqrt: the hardware seed plus one third-order step (within an ulp of IEEE).
  // (panel_factor.cuh: rolled, shifting, branch-free pivot loop -- instruction fetch, not arithmetic, is what a
  // once-per-launch kernel pays for)
  auto factor_sub = [&](int c0, double* xd) {
    const unsigned mask = factor_sub16<XDP, true>(S + (c0 + (lane & 15)) * LDS + c0, xd, cbuf, rdiag + c0 + (lane & 15), lane);
    if (mask && lane == 0 && *info == 0) *info = k0 + c0 + __ffs(mask);
  };
  // rank-16 update of the 16 x 16 sub-block t of the trailing lower triangle of step kb (t = 0: the next diagonal one)
  auto update_sub = [&](int kb, int t) {
    const int c0 = kb * SB;
    const int ii = lane & 15, jh = lane >> 4;
    int bi = 0;
    while ((bi + 1) * (bi + 2) / 2 <= t) ++bi;
    const int bj = t - bi * (bi + 1) / 2;
    const int ri = (kb + 1 + bi) * SB + ii;
    const int cj = (kb + 1 + bj) * SB + 8 * jh;
    double pr[SB];
#pragma unroll
    for (int l = 0; l < SB; ++l) pr[l] = S[ri * LDS + c0 + l];
    double acc[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const double* pj = S + (cj + jj) * LDS + c0;
      double sum = 0.0;
#pragma unroll
      for (int l = 0; l < SB; ++l) sum += pr[l] * pj[l];
      acc[jj] = sum;
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj)
      if (cj + jj <= ri) S[ri * LDS + cj + jj] -= acc[jj];
  };

  // rows below sub-block kb:  P[r][c] = sum_{l <= c} S[r][c0 + l] X[c][l]  for `ncol` columns per thread starting at
  // column cq; the threads of a row sit in one warp: all reads of the row precede its writes
  auto solve_rows = [&](int kb, const double* xd, int r, int cq, auto ncol_tag) {
    constexpr int NC = decltype(ncol_tag)::value;
    const int c0 = kb * SB;
    const bool act = r < NB;
    double out[NC];
    if (act) {
      double row[SB];
#pragma unroll
      for (int l = 0; l < SB; ++l) row[l] = S[r * LDS + c0 + l];
#pragma unroll
      for (int e = 0; e < NC; ++e) {
        const double* xr = xd + (cq + e) * XDP;
        double sum = 0.0;
#pragma unroll
        for (int l = 0; l < SB; ++l) sum += row[l] * xr[l];
        out[e] = sum;
      }
    }
    __syncwarp();
    if (act) {
#pragma unroll
      for (int e = 0; e < NC; ++e) S[r * LDS + c0 + cq + e] = out[e];
    }
  };
  // Block columns [kb_lo, kb_hi) of L (rows on and below the diagonal sub-block) and their inverted diagonal
  // sub-blocks, written by `nthr` threads (this one is number t).  Used by the three warps that idle during a phase --
  // block column kb - 1 is final once phase kb - 1 has ended -- and by everybody for the last two after the loop, so
  // that the kernel does not end on 80 KB of stores from one SM (300 KB with the packed copies of the distributed
  // factorisation).  Only the diagonal sub-blocks of the inverse are written: nothing reads the rest of a Linv slot
  // before linv_complete_kernel has rebuilt it (the panel TRSM and linv_complete_kernel use the sub-blocks alone), and
  // nobody reads the packed diagonal block above its diagonal sub-blocks (panel_unpack2_kernel copies c <= r).
  const bool early_store = !full_inverse;
  auto store_cols = [&](int kb_lo, int kb_hi, int t, int nthr) {
    for (int kb = kb_lo; kb < kb_hi; ++kb) {
      const int c0 = kb * SB;
      for (int e = t; e < (NB - c0) * SB; e += nthr) {
        const int r = c0 + (e >> 4), c = c0 + (e & 15);
        const bool in = r < nb && c <= r;
        const double v = in ? S[r * LDS + c] : 0.0;
        if (in) blk[(int64_t)r * ldm + c] = v;
        if (pack_lkk) pack_lkk[r * NB + c] = v;  // packed send buffer: what lies above the diagonal SUB-blocks is never read
      }
      for (int e = t; e < SB * SB; e += nthr) {
        const int i = c0 + (e >> 4), c = c0 + (e & 15);
        double v = 0.0;
        if (i < nb && c <= i) v = (c == i) ? rdiag[i] : S[c * LDS + i];
        Linv[i * NB + c] = v;
        if (pack_linv) pack_linv[i * NB + c] = v;
      }
    }
  };

  // Schedule.  Sub-block kb is factored by warp 0 alone (the serial chain of the panel); everything else is arranged
  // so that warp 0 never waits for more than it needs:
  //   prologue   warp 0: load + factor sub-block 0        | warps 1..15: load the rest of the block
  //   phase kb   warp 0: the 16 rows of block row kb + 1 times inv(L_kb)^T -> arrive at barrier 1 -> rank-16 update of
  //              the next diagonal sub-block -> factor it (into the other Xd buffer)
  //              warps 4, 8, 12: write block column kb - 1 of L and its inverted sub-block to global memory
  //              12 worker warps (not 4, 8, 12: those share warp 0's sub-partition): the rows of block rows >= kb + 2
  //              -> wait at barrier 1 (they need warp 0's rows) -> rank-16 update of the rest of the trailing triangle
  //   one CTA barrier per phase.  Same arithmetic per entry as the plain right-looking loop (every sub-block sees the
  //   same updates in the same order).  factor_sub has ONE call site (one copy of its code to fetch).
  for (int kb = -1; kb + 1 < NSB; ++kb) {
    double* xd_next = Xd + ((kb + 1) & 1) * (SB * XDP);
    if (warp == 0) {
      if (kb >= 0) {
        solve_rows(kb, Xd + (kb & 1) * (SB * XDP), (kb + 1) * SB + (lane & 15), 8 * (lane >> 4), std::integral_constant<int, 8>{});
        __threadfence_block();
        asm volatile("bar.arrive 1, 416;" ::: "memory");  // the 12 workers wait for these rows
        P2_T(2 + 4 * kb);
        __syncwarp();
        update_sub(kb, 0);
        __syncwarp();
        P2_T(3 + 4 * kb);
      }
      factor_sub((kb + 1) * SB, xd_next);
      if (kb >= 0) P2_T(4 + 4 * kb);
    } else if (kb >= 0 && (warp & 3) != 0) {
      // ---- 12 workers: the warps that share warp 0's SM sub-partition (4, 8, 12) are not among them -- their FP64 and
      // shared-memory instructions queued in front of the serial chain's and stretched it by up to 60 %.  (Splitting
      // warp 0's rows and diagonal update over those three as well was measured: no faster, both are bound by the
      // latency of their 16-term sums, and the extra code is fetched cold on every launch.)
      const int w = (warp >> 2) * 3 + (warp & 3) - 1, wt = w * 32 + lane;
      solve_rows(kb, Xd + (kb & 1) * (SB * XDP), (kb + 2) * SB + (wt >> 2), 4 * (wt & 3), std::integral_constant<int, 4>{});
      asm volatile("bar.sync 1, 416;" ::: "memory");
      const int nrem = NSB - 1 - kb;
      const int T = nrem * (nrem + 1) / 2;
      for (int t = 1 + w; t < T; t += 12) update_sub(kb, t);
    } else if (kb >= 1 && early_store) {
      store_cols(kb - 1, kb, (warp >> 2) * 32 - 32 + lane, 96);  // warps 4, 8, 12: block column kb - 1 is final since the last barrier
    }
    __syncthreads();  // inv(L_dd) of sub-block kb + 1 is in its Xd buffer; every update of step kb has landed
    if (kb >= 0) P2_T(5 + 4 * kb); else P2_T(1);
  }
  P2_T(35);

  // ---- L back to Mat (and, for the distributed factorisation, into the packed send buffer: dense 128 x 128, zero
  // above the diagonal and beyond a ragged block)
  if (early_store) {
    store_cols(NSB - 2, NSB, tid, 512);  // the last two block columns; the others left during the phases
  } else {
    for (int idx = tid; idx < NB * NB; idx += 512) {
      const int r = idx >> 7, c = idx & (NB - 1);
      const bool in = r < nb && c <= r;
      if (in) blk[(int64_t)r * ldm + c] = S[r * LDS + c];
      if (pack_lkk) pack_lkk[idx] = in ? S[r * LDS + c] : 0.0;
    }
    // ---- X to Linv.  Off-diagonal blocks of X: only the full-inverse consumers need them (trsm_impl 2); the default
    // TRSM uses the 16 x 16 diagonal inverses alone and the solves complete the inverses off the critical path
    // (linv_complete_kernel reads nothing but the diagonal sub-blocks), so the early-store path writes only those
    // 8 x 256 entries and the rest of the 128 x 128 slot keeps whatever the previous factorisation left there.
    if (full_inverse) complete_block_inverse(S, rdiag, Tw, tid);
    for (int idx = tid; idx < NB * NB; idx += 512) {
      const int i = idx >> 7, c = idx & (NB - 1);
      double v = 0.0;
      if (i < nb && c <= i) v = (c == i) ? rdiag[i] : S[c * LDS + i];
      Linv[idx] = v;
      if (pack_linv) pack_linv[idx] = v;
    }
  }
#ifdef LPB_POTF2_PROF
  P2_T(36);
  if (tid == 0 && k0 == 1024) {
    printf("potf2 k0=%d: load+factor0 %lld |", k0, p2t[1] - p2t[0]);
    for (int kb = 0; kb + 1 < NSB; ++kb)
      printf(" kb%d: rows16 %lld upd0 %lld factor %lld sync %lld |", kb, p2t[2 + 4 * kb] - p2t[1 + 4 * kb],
             p2t[3 + 4 * kb] - p2t[2 + 4 * kb], p2t[4 + 4 * kb] - p2t[3 + 4 * kb], p2t[5 + 4 * kb] - p2t[4 + 4 * kb]);
    printf(" stores %lld total %lld\n", p2t[36] - p2t[35], p2t[36] - p2t[0]);
  }
#endif
}

// ------------------------------------------------------------------ TRSM: X L_kk^T = B, 64 rows per CTA
// Warp w owns rows w, w+8, ... (8 rows); no cross-warp dependency, so only __syncwarp inside.
__global__ void __launch_bounds__(256)
trsm_kernel(double* __restrict__ Mat, int64_t ldm, int k0, int nb, int m) {
  extern __shared__ double sm[];
  double* SL = sm;              // NB * LDS : L_kk (lower incl. diag)
  double* SB = sm + NB * LDS;   // TRSM_ROWS * LDS : panel rows
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int r0 = k0 + nb + blockIdx.x * TRSM_ROWS;
  int nr = m - r0;
  if (nr > TRSM_ROWS) nr = TRSM_ROWS;
  const double* lkk = Mat + (int64_t)k0 * ldm + k0;
  for (int i = ty; i < nb; i += 8)
    for (int k = tx; k <= i; k += 32) SL[i * LDS + k] = lkk[(int64_t)i * ldm + k];
  double* pan = Mat + (int64_t)r0 * ldm + k0;
  for (int r = ty; r < nr; r += 8)
    for (int k = tx; k < nb; k += 32) SB[r * LDS + k] = pan[(int64_t)r * ldm + k];
  __syncthreads();

  for (int l = 0; l < nb; ++l) {
    const double dl = SL[l * LDS + l];
    double xl[TRSM_ROWS / 8];
#pragma unroll
    for (int rr = 0; rr < TRSM_ROWS / 8; ++rr) {
      const int r = ty + rr * 8;
      xl[rr] = (r < nr) ? SB[r * LDS + l] / dl : 0.0;
    }
    __syncwarp();
#pragma unroll
    for (int rr = 0; rr < TRSM_ROWS / 8; ++rr) {
      const int r = ty + rr * 8;
      if (r < nr) {
        for (int j = l + 1 + tx; j < nb; j += 32) SB[r * LDS + j] -= xl[rr] * SL[j * LDS + l];
        if (tx == 0) SB[r * LDS + l] = xl[rr];
      }
    }
    __syncwarp();
  }
  __syncthreads();
  for (int r = ty; r < nr; r += 8)
    for (int k = tx; k < nb; k += 32) pan[(int64_t)r * ldm + k] = SB[r * LDS + k];
}

// ------------------------------------------------------------------ panel TRSM, blocked substitution
// X L_kk^T = P for 64 panel rows per CTA.  L_kk is applied as 8 sub-block steps of 16 columns:
//     P[:, s] = T[:, s] inv(L_ss)^T ,   T[:, j] -= P[:, s] L[j][s]^T  for the later sub-blocks j,
// i.e. substitution between the 16-column sub-blocks and a 16 x 16 explicit inverse (the diagonal
// blocks of Linv, by-products of potf2_inv_kernel) inside each.  A GEMM with the full 128 x 128 inverse
// is faster on paper but NOT backward stable: late in the interior-point iteration, when M is very
// ill-conditioned, it perturbed d_tau enough to cost 3-4 extra iterations against the CPU oracle
// (C3: 28 instead of 24); with this form the trajectory follows the oracle's.  Plain DFMA: the whole
// TRSM is 128 m flop per panel column, ~1 % of the factorisation.
constexpr int TBR = 64;            // panel rows per CTA
constexpr int LDP = NB + 2;        // even pitch: 16-byte aligned double2 rows, conflict-free for 4-row groups
constexpr size_t kTrsmBlockedSmem = (size_t)((NB + TBR) * LDP + NSB * SB * XDP) * sizeof(double);

__global__ void __launch_bounds__(256)
trsm_blocked_kernel(double* __restrict__ Mat, int64_t ldm, int k0, int m, const double* __restrict__ Linv) {
  extern __shared__ __align__(16) double sm[];
  double* SL = sm;                      // NB * LDP : L_kk (lower triangle incl. diagonal)
  double* SP = SL + NB * LDP;           // TBR * LDP: panel rows
  double* SX = SP + TBR * LDP;          // NSB * SB * XDP: inv(L_ss), s = 0..7
  const int tid = threadIdx.x;
  const int64_t r0 = (int64_t)k0 + NB + (int64_t)blockIdx.x * TBR;
  const double* lkk = Mat + (int64_t)k0 * ldm + k0;
  for (int idx = tid; idx < NB * NB; idx += 256) {
    const int i = idx >> 7, c = idx & (NB - 1);
    if (c <= i) SL[i * LDP + c] = lkk[(int64_t)i * ldm + c];
  }
  for (int idx = tid; idx < NSB * SB * SB; idx += 256) {
    const int sblk = idx >> 8, i = (idx >> 4) & 15, l = idx & 15;
    SX[(sblk * SB + i) * XDP + l] = Linv[(sblk * SB + i) * NB + sblk * SB + l];
  }
  for (int idx = tid; idx < TBR * NB; idx += 256) {
    const int r = idx >> 7, c = idx & (NB - 1);
    SP[r * LDP + c] = (r0 + r < m) ? Mat[(r0 + r) * ldm + k0 + c] : 0.0;
  }
  __syncthreads();

  const int r = tid >> 2, q = tid & 3;  // the four threads of a row sit in one warp: only __syncwarp below
  double* prow = SP + r * LDP;
#pragma unroll 1
  for (int kb = 0; kb < NSB; ++kb) {
    const int c0 = kb * SB;
    double row[SB];
#pragma unroll
    for (int l = 0; l < SB; l += 2) {
      const double2 v = *reinterpret_cast<const double2*>(prow + c0 + l);
      row[l] = v.x;
      row[l + 1] = v.y;
    }
    double out[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const double* xr = SX + (kb * SB + 4 * q + e) * XDP;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int l = 0; l < SB; l += 2) {
        s0 += row[l] * xr[l];
        s1 += row[l + 1] * xr[l + 1];
      }
      out[e] = s0 + s1;
    }
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 4; ++e) prow[c0 + 4 * q + e] = out[e];
    __syncwarp();
#pragma unroll
    for (int l = 0; l < SB; l += 2) {
      const double2 v = *reinterpret_cast<const double2*>(prow + c0 + l);
      row[l] = v.x;
      row[l + 1] = v.y;
    }
    for (int j = c0 + SB + q; j < NB; j += 4) {
      const double* lj = SL + j * LDP + c0;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int l = 0; l < SB; l += 2) {
        const double2 v = *reinterpret_cast<const double2*>(lj + l);
        s0 += row[l] * v.x;
        s1 += row[l + 1] * v.y;
      }
      prow[j] -= s0 + s1;
    }
    __syncwarp();
  }
  __syncthreads();
  for (int idx = tid; idx < TBR * NB; idx += 256) {
    const int rr = idx >> 7, c = idx & (NB - 1);
    if (r0 + rr < m) Mat[(r0 + rr) * ldm + k0 + c] = SP[rr * LDP + c];
  }
}

// ------------------------------------------------------------------ panel TRSM, blocked substitution on the DMMA pipe
// The same algorithm as trsm_blocked_kernel (substitution between the 16-column sub-blocks, the explicit
// 16 x 16 inverses inside them) with every product on mma.sync.m8n8k4.f64.  A warp owns 8 panel rows and keeps
// the whole 8 x 128 slab in registers as 16 accumulator fragments c[q] (columns 8q .. 8q+7): sub-block step s
//     X_s  = T_s inv(L_ss)^T            : 2 fragments x 4 k-groups, B-fragments from the staged inverses
//     T_q -= X_s L[q-block][s-block]^T  : every later fragment q, 4 k-groups, B-fragments from the staged L_kk
// The A-operands (T_s, then -X_s) are accumulator fragments re-shaped by two quad shuffles per k-group
// (accumulator: thread t holds columns 2t, 2t+1 of row g; A-fragment: thread t holds column t).  288 DMMAs per
// warp instead of ~9000 dependent DFMAs per thread: the panel solve drops from ~44 us to a few us per wave.
// Shared pitches are = 4 mod 16 doubles so the 64-bit B-fragment loads (row g, column t) are conflict-free.
constexpr int TLP = NB + 4;   // 132
constexpr int TXP = SB + 4;   // 20
constexpr size_t kTrsmDmmaSmem = (size_t)(NB * TLP + NSB * SB * TXP) * sizeof(double);

__device__ __forceinline__ void dmma884(double2& c, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
      : "+d"(c.x), "+d"(c.y)
      : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256)
trsm_dmma_blocked_kernel(double* __restrict__ Mat, int64_t ldm, int k0, int m, const double* __restrict__ Linv,
                         double* __restrict__ pack = nullptr) {
  // pack != null (distributed factorisation): the solved rows also go into the packed send buffer of the panel --
  // 128 doubles per row, row (r - k0) at buffer row (r - k0), except that block k+1 (the rows the next owner needs
  // first) sits at buffer rows 0..127 and the diagonal block at rows 128..255 (written by potf2_inv_kernel).
  extern __shared__ __align__(16) double sm[];
  double* SL = sm;               // NB * TLP : L_kk (lower triangle incl. diagonal)
  double* SX = SL + NB * TLP;    // NSB * SB * TXP : inv(L_ss), s = 0..7 (zero above their diagonals)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const unsigned full = 0xffffffffu;
  const double* lkk = Mat + (int64_t)k0 * ldm + k0;
  for (int idx = tid; idx < NB * NB / 2; idx += 256) {
    const int i = idx >> 6, c = (idx & 63) * 2;
    if (c <= i)  // c + 1 may be i + 1: one entry above the diagonal, never read
      *reinterpret_cast<double2*>(SL + i * TLP + c) = __ldg(reinterpret_cast<const double2*>(lkk + (int64_t)i * ldm + c));
  }
  for (int idx = tid; idx < NSB * SB * SB; idx += 256) {
    const int sblk = idx >> 8, i = (idx >> 4) & 15, l = idx & 15;
    SX[(sblk * SB + i) * TXP + l] = __ldg(Linv + (sblk * SB + i) * NB + sblk * SB + l);
  }
  const int64_t r = (int64_t)k0 + NB + (int64_t)blockIdx.x * TBR + warp * 8 + g;
  double* prow = Mat + r * ldm + k0 + 2 * t;
  double2 c[16];
#pragma unroll
  for (int q = 0; q < 16; ++q)
    c[q] = (r < m) ? *reinterpret_cast<const double2*>(prow + 8 * q) : make_double2(0.0, 0.0);
  __syncthreads();

  const int src_lo = (lane & ~3) | (t >> 1);  // quad lane holding column (t) of the first half of a fragment
  const int src_hi = src_lo | 2;              // ... of the second half
  const bool odd = (t & 1) != 0;
#pragma unroll
  for (int s = 0; s < NSB; ++s) {
    // A-fragments of T_s: k-group kk covers columns 16 s + 4 kk + t
    double a[4];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const double2 v = c[2 * s + (kk >> 1)];
      const int src = (kk & 1) ? src_hi : src_lo;
      const double vx = __shfl_sync(full, v.x, src), vy = __shfl_sync(full, v.y, src);
      a[kk] = odd ? vy : vx;
    }
    double2 x[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      x[h] = make_double2(0.0, 0.0);
      const double* xb = SX + (s * SB + 8 * h + g) * TXP + t;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) dmma884(x[h], a[kk], xb[4 * kk]);
      c[2 * s + h] = x[h];
    }
    if (s + 1 < NSB) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const double2 v = x[kk >> 1];
        const int src = (kk & 1) ? src_hi : src_lo;
        const double vx = __shfl_sync(full, v.x, src), vy = __shfl_sync(full, v.y, src);
        a[kk] = -(odd ? vy : vx);
      }
#pragma unroll
      for (int q = 2 * s + 2; q < 16; ++q) {
        const double* lb = SL + (8 * q + g) * TLP + s * SB + t;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) dmma884(c[q], a[kk], lb[4 * kk]);
      }
    }
  }
  if (r < m) {
#pragma unroll
    for (int q = 0; q < 16; ++q) *reinterpret_cast<double2*>(prow + 8 * q) = c[q];
    if (pack) {
      int64_t br = r - k0;
      if (br < 2 * NB) br -= NB;
      double* brow = pack + br * NB + 2 * t;
#pragma unroll
      for (int q = 0; q < 16; ++q) *reinterpret_cast<double2*>(brow + 8 * q) = c[q];
    }
  }
}

// ------------------------------------------------------------------ K3: diagonal-block solve (one CTA)
// UPPER == false: L_kk w = b (forward).  UPPER == true: L_kk^T x = b (backward).
// Thread i owns row i of the block; reciprocal diagonals keep the division off the critical path.
__device__ __forceinline__ void bar_sync_128() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <bool UPPER, int NRHS>
__global__ void __launch_bounds__(512)
trsv_diag_kernel(const double* __restrict__ L, int64_t ldm, int k0, int nb, double* __restrict__ B, int64_t m) {
  extern __shared__ double sm[];
  double* SL = sm;                 // NB * LDS
  double* rdiag = SL + NB * LDS;   // NB
  double* sx = rdiag + NB;         // NRHS * NB
  const int i = threadIdx.x;
  const double* lkk = L + (int64_t)k0 * ldm + k0;
  // all 16 warps stream the lower triangle in (coalesced rows), then 4 warps run the substitution
#pragma unroll 4
  for (int idx = i; idx < nb * NB; idx += 512) {
    const int r = idx >> 7, c = idx & (NB - 1);
    if (c <= r) SL[r * LDS + c] = __ldg(lkk + (int64_t)r * ldm + c);
  }
  __syncthreads();
  if (i >= NB) return;
  if (i < nb) rdiag[i] = 1.0 / SL[i * LDS + i];
  double b0 = 0.0, b1 = 0.0;
  if (i < nb) {
    b0 = B[k0 + i];
    if (NRHS == 2) b1 = B[m + k0 + i];
  }
  bar_sync_128();
  if (!UPPER) {
    for (int l = 0; l < nb; ++l) {
      if (i == l) {
        sx[l] = b0 = b0 * rdiag[l];
        if (NRHS == 2) sx[NB + l] = b1 = b1 * rdiag[l];
      }
      bar_sync_128();
      if (i > l && i < nb) {
        const double lil = SL[i * LDS + l];
        b0 -= sx[l] * lil;
        if (NRHS == 2) b1 -= sx[NB + l] * lil;
      }
    }
  } else {
    for (int l = nb - 1; l >= 0; --l) {
      if (i == l) {
        sx[l] = b0 = b0 * rdiag[l];
        if (NRHS == 2) sx[NB + l] = b1 = b1 * rdiag[l];
      }
      bar_sync_128();
      if (i < l) {
        const double lli = SL[l * LDS + i];
        b0 -= sx[l] * lli;
        if (NRHS == 2) b1 -= sx[NB + l] * lli;
      }
    }
  }
  if (i < nb) {
    B[k0 + i] = b0;
    if (NRHS == 2) B[m + k0 + i] = b1;
  }
}

// forward update: B[r] -= sum_c L[r][k0+c] x[c] for all rows r >= k0+nb.  Warp per row, 8 rows per warp.
template <int NRHS>
__global__ void __launch_bounds__(256)
trsv_update_fwd_kernel(const double* __restrict__ L, int64_t ldm, int k0, int nb, double* __restrict__ B, int64_t m) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // x block in registers: lane holds columns lane, lane+32, lane+64, lane+96
  double x0[4], x1[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int c = lane + 32 * q;
    x0[q] = c < nb ? B[k0 + c] : 0.0;
    x1[q] = (NRHS == 2 && c < nb) ? B[m + k0 + c] : 0.0;
  }
  const int64_t rbase = (int64_t)k0 + nb + ((int64_t)blockIdx.x * 8 + warp) * 8;
#pragma unroll 2
  for (int rr = 0; rr < 8; ++rr) {
    const int64_t r = rbase + rr;
    if (r >= m) break;
    const double* lr = L + r * ldm + k0;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = lane + 32 * q;
      const double a = c < nb ? __ldg(lr + c) : 0.0;
      s0 += a * x0[q];
      if (NRHS == 2) s1 += a * x1[q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      if (NRHS == 2) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) {
      B[r] -= s0;
      if (NRHS == 2) B[m + r] -= s1;
    }
  }
}

// backward update: B[c] -= sum_r L[k0+r][c] x[r] for all columns c < k0.
// Block = 128 columns x 4 row groups; rows of L are contiguous so every load is coalesced.
template <int NRHS>
__global__ void __launch_bounds__(512)
trsv_update_bwd_kernel(const double* __restrict__ L, int64_t ldm, int k0, int nb, double* __restrict__ B, int64_t m) {
  __shared__ double sx[NRHS][NB];
  __shared__ double part[NRHS][4][128];
  const int tx = threadIdx.x, ty = threadIdx.y;  // 128 x 4
  for (int i = ty * 128 + tx; i < nb; i += 512) {
    sx[0][i] = B[k0 + i];
    if (NRHS == 2) sx[1][i] = B[m + k0 + i];
  }
  __syncthreads();
  const int c = blockIdx.x * 128 + tx;
  double s0 = 0.0, s1 = 0.0;
  if (c < k0) {
    const double* lp = L + (int64_t)k0 * ldm + c;
#pragma unroll 8
    for (int r = ty; r < nb; r += 4) {
      const double a = __ldg(lp + (int64_t)r * ldm);
      s0 += a * sx[0][r];
      if (NRHS == 2) s1 += a * sx[1][r];
    }
  }
  part[0][ty][tx] = s0;
  if (NRHS == 2) part[1][ty][tx] = s1;
  __syncthreads();
  if (ty == 0 && c < k0) {
    B[c] -= (part[0][0][tx] + part[0][1][tx]) + (part[0][2][tx] + part[0][3][tx]);
    if (NRHS == 2) B[m + c] -= (part[1][0][tx] + part[1][1][tx]) + (part[1][2][tx] + part[1][3][tx]);
  }
}

// ------------------------------------------------------------------ K3 fast path: fused solve steps
// Forward step k (L w = b):  every CTA computes w_k = inv(L_kk) b_k (Linv block out of L2); CTA 0 also
// publishes it to Y; CTA j then applies b_i -= L[i][k-block] w_k to its own 128 rows i (blockIdx.x >= 1).
// Grid = 1 + number of 128-row blocks below block k.  B is the running right-hand side, Y the result.
template <int NRHS>
__global__ void __launch_bounds__(256)
solve_fwd_step_kernel(const double* __restrict__ L, int64_t ldm, const double* __restrict__ Linv, int k0, int nb,
                      double* __restrict__ B, double* __restrict__ Y, int64_t m) {
  __shared__ double sb[NRHS][NB];
  __shared__ double sw[NRHS][NB];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < NB; i += 256) {
    sb[0][i] = i < nb ? B[k0 + i] : 0.0;
    if (NRHS == 2) sb[1][i] = i < nb ? B[m + k0 + i] : 0.0;
  }
  __syncthreads();
  // w = Linv * b : warp per row (16 rows per warp), lanes over the 128 columns
  for (int r = warp; r < NB; r += 8) {
    const double* lr = Linv + r * NB;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = lane + 32 * q;
      const double a = c <= r ? __ldg(lr + c) : 0.0;
      s0 += a * sb[0][c];
      if (NRHS == 2) s1 += a * sb[1][c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      if (NRHS == 2) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) {
      sw[0][r] = s0;
      if (NRHS == 2) sw[1][r] = s1;
    }
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    for (int i = tid; i < nb; i += 256) {
      Y[k0 + i] = sw[0][i];
      if (NRHS == 2) Y[m + k0 + i] = sw[1][i];
    }
    return;
  }
  // rank-nb update of this CTA's 128 rows: warp per row
  const int64_t rbase = (int64_t)k0 + nb + (int64_t)(blockIdx.x - 1) * NB;
  double x0[4], x1[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    x0[q] = sw[0][lane + 32 * q];
    x1[q] = NRHS == 2 ? sw[1][lane + 32 * q] : 0.0;
  }
#pragma unroll 4
  for (int rr = warp; rr < NB; rr += 8) {
    const int64_t r = rbase + rr;
    if (r >= m) break;
    const double* lr = L + r * ldm + k0;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = lane + 32 * q;
      const double a = c < nb ? __ldg(lr + c) : 0.0;
      s0 += a * x0[q];
      if (NRHS == 2) s1 += a * x1[q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      if (NRHS == 2) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) {
      B[r] -= s0;
      if (NRHS == 2) B[m + r] -= s1;
    }
  }
}

// Backward step k (L^T x = w):  every CTA computes x_k = inv(L_kk)^T y_k; CTA 0 publishes it to X; CTA j
// applies y_c -= sum_r L[k0+r][c] x_k[r] to its own 128 columns c of block j-1 (< k).
template <int NRHS>
__global__ void __launch_bounds__(256)
solve_bwd_step_kernel(const double* __restrict__ L, int64_t ldm, const double* __restrict__ Linv, int k0, int nb,
                      double* __restrict__ Y, double* __restrict__ X, int64_t m) {
  __shared__ double sy[NRHS][NB];
  __shared__ double sxk[NRHS][NB];
  __shared__ double part[NRHS][2][NB];
  const int tid = threadIdx.x;
  const int c = tid & (NB - 1), h = tid >> 7;  // 128 columns x 2 row halves
  for (int i = tid; i < NB; i += 256) {
    sy[0][i] = i < nb ? Y[k0 + i] : 0.0;
    if (NRHS == 2) sy[1][i] = i < nb ? Y[m + k0 + i] : 0.0;
  }
  __syncthreads();
  {  // x[c] = sum_{r >= c} Linv[r][c] y[r]  (rows contiguous: coalesced across c)
    double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
    for (int r = h * 64; r < h * 64 + 64; ++r) {
      if (r >= c && r < nb) {
        const double a = __ldg(Linv + r * NB + c);
        s0 += a * sy[0][r];
        if (NRHS == 2) s1 += a * sy[1][r];
      }
    }
    part[0][h][c] = s0;
    if (NRHS == 2) part[1][h][c] = s1;
  }
  __syncthreads();
  if (h == 0) {
    sxk[0][c] = part[0][0][c] + part[0][1][c];
    if (NRHS == 2) sxk[1][c] = part[1][0][c] + part[1][1][c];
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    if (h == 0 && c < nb) {
      X[k0 + c] = sxk[0][c];
      if (NRHS == 2) X[m + k0 + c] = sxk[1][c];
    }
    return;
  }
  const int64_t col = (int64_t)(blockIdx.x - 1) * NB + c;  // < k0 always (k0 is a multiple of NB)
  double s0 = 0.0, s1 = 0.0;
  const double* lp = L + (int64_t)k0 * ldm + col;
  const int r_end = (h * 64 + 64) < nb ? (h * 64 + 64) : nb;
#pragma unroll 8
  for (int r = h * 64; r < r_end; ++r) {
    const double a = __ldg(lp + (int64_t)r * ldm);
    s0 += a * sxk[0][r];
    if (NRHS == 2) s1 += a * sxk[1][r];
  }
  part[0][h][c] = s0;
  if (NRHS == 2) part[1][h][c] = s1;
  __syncthreads();
  if (h == 0) {
    Y[col] -= part[0][0][c] + part[0][1][c];
    if (NRHS == 2) Y[m + col] -= part[1][0][c] + part[1][1][c];
  }
}

// ------------------------------------------------------------------ K3 v3: pipelined solve, ONE launch
// L L^T X = B for 1 or 2 right-hand sides in a single cooperative kernel.  Block row j (128 rows) is
// owned by CTA j mod gridDim.x; CTAs exchange the 128-entry solution blocks through global memory and
// per-block flags (release / acquire at gpu scope), so the only serial chain is
//     w_{j-1} published -> CTA j applies L[j][j-1] w_{j-1}, multiplies by inv(L_jj), publishes w_j
// (~3 us per block instead of one kernel launch per block), while every other block product
// L[j][k] w_k, k < j-1, is consumed as soon as w_k exists -- each CTA runs ahead of the chain.
// Warps are autonomous inside the sweep over k (each polls the flag itself: no CTA barrier per block).
// Forward:  warp = 16 rows, lane = 4 columns; per-lane partial sums are kept over ALL k and folded
//           across lanes once per block row (multi-value butterfly: 16 shuffles for 16 rows).
// Backward: the same loads of L[k][j] (rows of block k, columns of block j), accumulated per column;
//           the 8 warps are folded through shared memory.
// Co-residency of all CTAs is what makes the flag waits safe: the launch is cooperative and the grid
// is capped at the number of CTAs the device can hold.
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Bounded without trapping (a trap leaves a sticky error that kills every context of the process): after ~20 s
// of wall-clock polling the waiter raises the context's fault word and carries on; every other poller sees the
// word and leaves too; the host reports LPB_ERR_CUDA at its next scalar fetch.
// `relaxed` pollers (blocks that are not next in the chain) back off so that up to ~1000 warps
// spinning on one L2 line do not delay the release store they are waiting for.
__device__ __forceinline__ unsigned long long wall_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __noinline__ void wait_flag_slow(const int* flag, int epoch, bool relaxed, unsigned long long* fault) {
  const unsigned long long t0 = wall_ns();
  unsigned spins = 0;
  while (ld_acquire_gpu(flag) != epoch) {
    if (relaxed) __nanosleep(256);
    if ((++spins & 255u) == 0) {
      if (*reinterpret_cast<volatile unsigned long long*>(fault) != 0ull) return;
      if (wall_ns() - t0 > 20000000000ull) {
        atomicExch(fault, 1ull);
        return;
      }
    }
  }
}
__device__ __forceinline__ void wait_flag(const int* flag, int epoch, unsigned long long* fault, bool relaxed = false) {
  if (ld_acquire_gpu(flag) == epoch) return;
  wait_flag_slow(flag, epoch, relaxed, fault);
}
// Per-lane partials of 16 rows -> total of row (lane >> 1) in every lane (pairs hold copies).
__device__ __forceinline__ double fold16(double (&v)[16], int lane) {
  const unsigned full = 0xffffffffu;
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1) {
    const bool up = (lane & (2 * w)) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const double send = up ? v[i] : v[i + w];
      const double keep = up ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(full, send, 2 * w);
    }
  }
  return v[0] + __shfl_xor_sync(full, v[0], 1);
}

constexpr int kSolveThreads = 256;
constexpr size_t kSolveSmem = (size_t)(NB * LDP + NSB * SB * XDP + 2 * 2 * NB + 8 * 2 * NB) * sizeof(double);

// Stage L_jj (lower triangle, rows beyond m left out) and the 8 inverted 16 x 16 diagonal sub-blocks.
__device__ __forceinline__ void solve_stage_diag(const double* __restrict__ L, int64_t ldm, const double* __restrict__ Linv_j,
                                                 int64_t j0, int64_t m, double* sL, double* sX, int tid) {
  for (int idx = tid; idx < NB * NB / 2; idx += kSolveThreads) {
    const int i = idx >> 6, c = (idx & 63) * 2;
    if (c <= i) {  // c + 1 may be i + 1: one entry above the diagonal, never read
      double2 v = make_double2(0.0, 0.0);  // rows beyond m (ragged last block) are zero, not stale
      if (j0 + i < m) v = __ldg(reinterpret_cast<const double2*>(L + (j0 + i) * ldm + j0 + c));
      *reinterpret_cast<double2*>(sL + i * LDP + c) = v;
    }
  }
  for (int idx = tid; idx < NSB * SB * SB; idx += kSolveThreads) {
    const int sblk = idx >> 8, i = (idx >> 4) & 15, l = idx & 15;
    sX[(sblk * SB + i) * XDP + l] = __ldg(Linv_j + (sblk * SB + i) * NB + sblk * SB + l);
  }
}

// w = inv(L_jj) c, in place in sc (forward) -- blocked substitution over the 16-column sub-blocks with the
// 16 x 16 inverses inside (see trsm_blocked_kernel for why not one GEMV with the full inverse).
template <int NRHS>
__device__ __forceinline__ void solve_diag_fwd(const double* sL, const double* sX, double* sc, double* sw, int tid) {
  const int r = tid & (NB - 1), q = tid >> 7;
#pragma unroll 1
  for (int s = 0; s < NSB; ++s) {
    const int c0 = s * SB;
    if (tid < SB * NRHS) {
      const int i = tid & 15, qq = tid >> 4;
      const double* xr = sX + (c0 + i) * XDP;
      const double* cs = sc + qq * NB + c0;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int l = 0; l < SB; l += 2) {  // X is zero above its diagonal
        a0 += xr[l] * cs[l];
        a1 += xr[l + 1] * cs[l + 1];
      }
      sw[qq * NB + c0 + i] = a0 + a1;
    }
    __syncthreads();
    if (q < NRHS && r >= c0 + SB) {
      const double* lr = sL + r * LDP + c0;
      const double* ws = sw + q * NB + c0;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int l = 0; l < SB; l += 2) {
        a0 += lr[l] * ws[l];
        a1 += lr[l + 1] * ws[l + 1];
      }
      sc[q * NB + r] -= a0 + a1;
    }
    __syncthreads();
  }
}

// x = inv(L_jj)^T c (backward)
template <int NRHS>
__device__ __forceinline__ void solve_diag_bwd(const double* sL, const double* sX, double* sc, double* sw, int tid) {
  const int c = tid & (NB - 1), q = tid >> 7;
#pragma unroll 1
  for (int s = NSB - 1; s >= 0; --s) {
    const int c0 = s * SB;
    if (tid < SB * NRHS) {
      const int i = tid & 15, qq = tid >> 4;
      const double* cs = sc + qq * NB + c0;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int l = 0; l < SB; l += 2) {  // x_i = sum_l X[l][i] c_l ; X[l][i] = 0 for l < i
        a0 += sX[(c0 + l) * XDP + i] * cs[l];
        a1 += sX[(c0 + l + 1) * XDP + i] * cs[l + 1];
      }
      sw[qq * NB + c0 + i] = a0 + a1;
    }
    __syncthreads();
    if (q < NRHS && c < c0) {
      const double* xs = sw + q * NB + c0;
      double a0 = 0.0, a1 = 0.0;
#pragma unroll
      for (int l = 0; l < SB; l += 2) {
        a0 += sL[(c0 + l) * LDP + c] * xs[l];
        a1 += sL[(c0 + l + 1) * LDP + c] * xs[l + 1];
      }
      sc[q * NB + c] -= a0 + a1;
    }
    __syncthreads();
  }
}

template <int NRHS>
__global__ void __launch_bounds__(kSolveThreads, 1)
solve_pipelined_kernel(const double* __restrict__ L, int64_t ldm, const double* __restrict__ Linv_all, int64_t m,
                       int nblk, double* B, double* Y, int64_t ldy, int* flags_f, int* flags_b, int epoch,
                       unsigned long long* fault) {
  extern __shared__ __align__(16) double sm[];
  double* sL = sm;                        // 128 x LDP: L_jj
  double* sX = sL + NB * LDP;             // 8 x 16 x XDP: inv of its 16 x 16 diagonal sub-blocks
  double* sc = sX + NSB * SB * XDP;       // [2][128] right-hand side of the current block
  double* sw = sc + 2 * NB;               // [2][128] solution of the current block
  double* sred = sw + 2 * NB;             // [8 warps][2][128] backward cross-warp fold
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c4 = 4 * lane;                // this lane's 4 columns inside a 128-wide block

  // ============================================================ forward: L w = b
  for (int j = blockIdx.x; j < nblk; j += gridDim.x) {
    const int64_t r0 = (int64_t)j * NB;
    const int nbj = (int)((m - r0) < NB ? (m - r0) : NB);
    __syncthreads();  // the previous block row is done with sL / sX / sc / sw
    solve_stage_diag(L, ldm, Linv_all + (int64_t)j * NB * NB, r0, m, sL, sX, tid);  // latency hides behind the sweep
    double acc[NRHS][16];
#pragma unroll
    for (int q = 0; q < NRHS; ++q)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[q][i] = 0.0;
    const int64_t rw = r0 + warp * 16;  // first row of this warp
    for (int k = 0; k < j; ++k) {
      const int64_t k0 = (int64_t)k * NB;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        double2 la[8], lb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t r = rw + half * 8 + i;
          la[i] = lb[i] = make_double2(0.0, 0.0);
          if (r < m) {
            const double2* p = reinterpret_cast<const double2*>(L + r * ldm + k0 + c4);
            la[i] = __ldg(p);
            lb[i] = __ldg(p + 1);
          }
        }
        if (half == 0) wait_flag(flags_f + k, epoch, fault, k + 1 < j);
        double2 wa[NRHS], wb[NRHS];
#pragma unroll
        for (int q = 0; q < NRHS; ++q) {
          const double2* p = reinterpret_cast<const double2*>(Y + (int64_t)q * ldy + k0 + c4);
          wa[q] = __ldcg(p);
          wb[q] = __ldcg(p + 1);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int q = 0; q < NRHS; ++q)
            acc[q][half * 8 + i] += la[i].x * wa[q].x + la[i].y * wa[q].y + lb[i].x * wb[q].x + lb[i].y * wb[q].y;
      }
    }
    // fold across lanes; c = b - sum
#pragma unroll
    for (int q = 0; q < NRHS; ++q) {
      const double tot = fold16(acc[q], lane);
      const int rr = warp * 16 + (lane >> 1);
      if ((lane & 1) == 0) sc[q * NB + rr] = (rr < nbj) ? B[(int64_t)q * m + r0 + rr] - tot : 0.0;
    }
    __syncthreads();
    solve_diag_fwd<NRHS>(sL, sX, sc, sw, tid);
    if (tid < NRHS * NB) {
      const int q = tid >> 7, rr = tid & (NB - 1);
      if (rr < nbj) Y[(int64_t)q * ldy + r0 + rr] = sw[q * NB + rr];
    }
    __syncthreads();  // all of w_j written
    if (tid == 0) st_release_gpu(flags_f + j, epoch);  // release is cumulative over the barrier: no extra fence
  }

  // ============================================================ backward: L^T x = w
  for (int j = nblk - 1 - blockIdx.x; j >= 0; j -= gridDim.x) {
    const int64_t j0 = (int64_t)j * NB;
    const int nbj = (int)((m - j0) < NB ? (m - j0) : NB);
    __syncthreads();  // previous users of sL / sX / sc / sw / sred are done
    solve_stage_diag(L, ldm, Linv_all + (int64_t)j * NB * NB, j0, m, sL, sX, tid);
    double acc[NRHS][4];
#pragma unroll
    for (int q = 0; q < NRHS; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0;
    for (int k = nblk - 1; k > j; --k) {
      const int64_t k0 = (int64_t)k * NB;
      const int64_t rw = k0 + warp * 16;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        double2 la[8], lb[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t r = rw + half * 8 + i;
          la[i] = lb[i] = make_double2(0.0, 0.0);
          if (r < m) {
            const double2* p = reinterpret_cast<const double2*>(L + r * ldm + j0 + c4);
            la[i] = __ldg(p);
            lb[i] = __ldg(p + 1);
          }
        }
        if (half == 0) wait_flag(flags_b + k, epoch, fault, k - 1 > j);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t r = rw + half * 8 + i;
#pragma unroll
          for (int q = 0; q < NRHS; ++q) {
            const double xv = (r < m) ? __ldcg(B + (int64_t)q * m + r) : 0.0;
            acc[q][0] += la[i].x * xv;
            acc[q][1] += la[i].y * xv;
            acc[q][2] += lb[i].x * xv;
            acc[q][3] += lb[i].y * xv;
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NRHS; ++q)
      *reinterpret_cast<double2*>(sred + (warp * 2 + q) * NB + c4) = make_double2(acc[q][0], acc[q][1]),
      *reinterpret_cast<double2*>(sred + (warp * 2 + q) * NB + c4 + 2) = make_double2(acc[q][2], acc[q][3]);
    __syncthreads();
    if (tid < NRHS * NB) {  // c = w_j - sum over the 8 warps
      wait_flag(flags_f + j, epoch, fault);  // w_j may come from another CTA (forward owner j mod G)
      const int q = tid >> 7, c = tid & (NB - 1);
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += sred[(w * 2 + q) * NB + c];
      sc[q * NB + c] = (c < nbj) ? __ldcg(Y + (int64_t)q * ldy + j0 + c) - sum : 0.0;
    }
    __syncthreads();
    solve_diag_bwd<NRHS>(sL, sX, sc, sw, tid);
    if (tid < NRHS * NB) {
      const int q = tid >> 7, c = tid & (NB - 1);
      if (c < nbj) B[(int64_t)q * m + j0 + c] = sw[q * NB + c];
    }
    __syncthreads();
    if (tid == 0) st_release_gpu(flags_b + j, epoch);
  }
}

// ------------------------------------------------------------------ K3 v4: full block inverses + tagged hand-off
// linv_complete_kernel: one CTA per diagonal block, AFTER the factorisation (off its critical path, all blocks in
// parallel): completes X_j = inv(L_jj) from L_jj and the inverted 16 x 16 diagonal sub-blocks that potf2_inv left
// in Linv (dense 128 x 128 per block, zero above the diagonal and beyond a ragged last block).
constexpr size_t kLinvCompleteSmem = (size_t)(NB * LDS + NB + 16 * SB * TWP) * sizeof(double);

__global__ void __launch_bounds__(512)
linv_complete_kernel(const double* __restrict__ Mat, int64_t ldm, int64_t m, double* __restrict__ Linv_all) {
  extern __shared__ double S[];   // NB * LDS: L_jj below / on the diagonal, X^T above
  double* rdiag = S + NB * LDS;   // diagonal of X
  double* Tw = rdiag + NB;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int64_t j0 = (int64_t)blockIdx.x * NB;
  const int nb = (int)((m - j0) < NB ? (m - j0) : NB);
  double* Xj = Linv_all + (int64_t)blockIdx.x * NB * NB;
  const double* blk = Mat + j0 * ldm + j0;
  for (int idx = tid; idx < NB * NB; idx += 512) {
    const int r = idx >> 7, c = idx & (NB - 1);
    if (c <= r) {
      S[r * LDS + c] = (r < nb) ? blk[(int64_t)r * ldm + c] : (r == c ? 1.0 : 0.0);  // identity padding
      const bool same_sub = (r >> 4) == (c >> 4);
      const double xv = (r < nb) ? (same_sub ? Xj[idx] : 0.0) : (r == c ? 1.0 : 0.0);
      if (r == c) rdiag[r] = xv;
      else S[c * LDS + r] = xv;   // X^T above the diagonal; off-diagonal sub-blocks start at zero
    }
  }
  __syncthreads();
  complete_block_inverse(S, rdiag, Tw, tid);
  for (int idx = tid; idx < NB * NB; idx += 512) {
    const int i = idx >> 7, c = idx & (NB - 1);
    double v = 0.0;
    if (i < nb && c <= i) v = (c == i) ? rdiag[i] : S[c * LDS + i];
    Xj[idx] = v;
  }
}

// solve_ll_kernel: L L^T X = B for 1 or 2 right-hand sides in ONE cooperative launch, like solve_pipelined_kernel
// (block row j owned by CTA j mod grid, every off-chain block product consumed as soon as its input exists), with
// the serial chain  "w_{j-1} visible -> w_j visible"  cut to one L2 store + one L2 load + two 128 x 128 GEMVs:
//   * hand-off without flags or fences: every solution entry travels as a 16-byte word {lo, tag, hi, tag}
//     (tag = launch epoch) written with one st.volatile.v4 and polled with ld.volatile.v4 -- the LL scheme of
//     NCCL: it only needs 8-byte store atomicity, and data + validity arrive in the same transaction;
//   * w_j = X_j c_j with the FULL inverse X_j = inv(L_jj) staged in shared memory (one GEMV, two threads per
//     row) instead of 8 dependent sub-block steps with two CTA barriers each;
//   * the L block of the next product is in registers BEFORE its w is polled, so the chain never waits on HBM.
// A lost hand-off does not hang or trap: after ~20 s of polling the waiter raises the context's fault word, every
// other poller sees it and leaves, and the host reports LPB_ERR_CUDA at its next scalar fetch.
struct __align__(16) Tagged {
  uint32_t lo, tag0, hi, tag1;
};
__device__ __forceinline__ void st_tagged(Tagged* p, double v, uint32_t tag) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((uint32_t)__double2loint(v)), "r"(tag),
               "r"((uint32_t)__double2hiint(v)), "r"(tag)
               : "memory");
}
__device__ __forceinline__ bool ld_tagged_try(const Tagged* p, uint32_t tag, double* v) {
  uint32_t a, b, c, d;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
  *v = __hiloint2double((int)c, (int)a);
  return b == tag && d == tag;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __noinline__ double ld_tagged_slow(const Tagged* p, uint32_t tag, unsigned long long* fault, bool relaxed) {
  double v = 0.0;
  const unsigned long long t0 = global_ns();
  unsigned spins = 0;
  while (!ld_tagged_try(p, tag, &v)) {
    if (relaxed) __nanosleep(200);
    if ((++spins & 255u) == 0) {
      if (*reinterpret_cast<volatile unsigned long long*>(fault) != 0ull) break;
      if (global_ns() - t0 > 20000000000ull) {
        atomicExch(fault, 1ull);
        break;
      }
    }
  }
  return v;
}
__device__ __forceinline__ double ld_tagged(const Tagged* p, uint32_t tag, unsigned long long* fault, bool relaxed) {
  double v;
  if (ld_tagged_try(p, tag, &v)) return v;
  return ld_tagged_slow(p, tag, fault, relaxed);
}

// Poll NQ x K tagged words (K consecutive entries for each of NQ right-hand sides, `qstride` entries apart) until
// every tag matches.  All loads of a round are issued back to back and checked together: one L2 round trip per
// round, however many words (a load-check-branch per word would serialise the round trips -- measured: 8 words
// cost 4 us of a 5 us chain step).
template <int NQ, int K>
__device__ __forceinline__ void poll_tagged(const Tagged* p, int qstride, uint32_t tag, unsigned long long* fault,
                                            double (&v)[NQ][K]) {
  unsigned spins = 0;
  unsigned long long t0 = 0ull;
  for (;;) {
    uint32_t r[NQ][K][4];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int e = 0; e < K; ++e)
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[q][e][0]), "=r"(r[q][e][1]), "=r"(r[q][e][2]), "=r"(r[q][e][3])
                     : "l"(p + q * qstride + e)
                     : "memory");
    bool ok = true;
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int e = 0; e < K; ++e) {
        ok = ok && r[q][e][1] == tag && r[q][e][3] == tag;
        v[q][e] = __hiloint2double((int)r[q][e][2], (int)r[q][e][0]);
      }
    if (ok) return;
    if ((++spins & 63u) == 0) {
      if (t0 == 0ull) t0 = global_ns();
      if (*reinterpret_cast<volatile unsigned long long*>(fault) != 0ull) return;
      if (global_ns() - t0 > 20000000000ull) {
        atomicExch(fault, 1ull);
        return;
      }
    }
  }
}

constexpr int XP = NB + 2;  // 130 = 2 mod 16: two threads per row (even / odd columns) read X_j conflict-free
constexpr size_t kSolveLLSmem = (size_t)(NB * XP + 2 * NB + 8 * 2 * NB) * sizeof(double);

// Stage X_j (TRANS: its transpose) into shared memory, pitch XP, zero above the diagonal.
template <bool TRANS>
__device__ __forceinline__ void stage_inverse(const double* __restrict__ Xj, double* sX, int tid) {
  for (int idx = tid; idx < NB * NB / 2; idx += kSolveThreads) {
    const int i = idx >> 6, c = (idx & 63) * 2;
    const double2 v = __ldg(reinterpret_cast<const double2*>(Xj + i * NB + c));
    if (TRANS) {
      sX[c * XP + i] = v.x;
      sX[(c + 1) * XP + i] = v.y;
    } else {
      sX[i * XP + c] = v.x;
      sX[i * XP + c + 1] = v.y;
    }
  }
}

// y = T c for the staged 128 x 128 triangular matrix T (X_j: lower, or X_j^T: upper): thread (row = tid >> 1,
// h = tid & 1) takes the columns = h mod 2; four independent chains per right-hand side; the two halves meet in one
// shuffle.  A warp owns 16 consecutive rows and only walks the 16-column groups that hold non-zeros for them
// (reading X_j out of shared memory is what bounds this step: 128 bytes per clock per SM).
template <int NRHS, bool UPPER>
__device__ __forceinline__ void block_gemv(const double* sX, const double* sc, int tid, double (&out)[NRHS]) {
  const int r = tid >> 1, h = tid & 1, wrow = (tid >> 5) * 16;
  const double* xr = sX + r * XP + h;
  double s[NRHS][4];
#pragma unroll
  for (int q = 0; q < NRHS; ++q) s[q][0] = s[q][1] = s[q][2] = s[q][3] = 0.0;
  const int i_lo = UPPER ? wrow / 2 : 0, i_hi = UPPER ? NB / 2 : (wrow + 16) / 2;
#pragma unroll 2
  for (int i = i_lo; i < i_hi; i += 4) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const double xv = xr[2 * (i + e)];
#pragma unroll
      for (int q = 0; q < NRHS; ++q) s[q][e] = fma(xv, sc[q * NB + 2 * (i + e) + h], s[q][e]);
    }
  }
#pragma unroll
  for (int q = 0; q < NRHS; ++q) {
    const double t = (s[q][0] + s[q][1]) + (s[q][2] + s[q][3]);
    out[q] = t + __shfl_xor_sync(0xffffffffu, t, 1);
  }
}

template <int NRHS>
__global__ void __launch_bounds__(kSolveThreads, 1)
solve_ll_kernel(const double* __restrict__ L, int64_t ldm, const double* __restrict__ Xall, int64_t m, int nblk,
                double* B, Tagged* Wt, Tagged* Xt, uint32_t epoch, unsigned long long* fault) {
  extern __shared__ __align__(16) double sm[];
  double* sX = sm;                 // 128 x XP: X_j (forward) / X_j^T (backward)
  double* sc = sX + NB * XP;       // [2][128] right-hand side of the current block
  double* sred = sc + 2 * NB;      // [8 warps][2][128] backward cross-warp fold
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c4 = 4 * lane;         // this lane's 4 columns inside a 128-wide block
  const unsigned full = 0xffffffffu;

  // ============================================================ forward: L w = b
  for (int j = blockIdx.x; j < nblk; j += gridDim.x) {
    const int64_t r0 = (int64_t)j * NB;
    const int nbj = (int)((m - r0) < NB ? (m - r0) : NB);
    __syncthreads();  // the previous block row is done with sX / sc
    stage_inverse<false>(Xall + (int64_t)j * NB * NB, sX, tid);
    double acc[NRHS][16];
#pragma unroll
    for (int q = 0; q < NRHS; ++q)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[q][i] = 0.0;
    const int64_t rw = r0 + warp * 16;  // first row of this warp
    double2 la[16], lb[16];
    auto load_block = [&](int k) {
      const double* base = L + (int64_t)k * NB + c4;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int64_t r = rw + i;
        la[i] = lb[i] = make_double2(0.0, 0.0);
        if (r < m) {
          const double2* p = reinterpret_cast<const double2*>(base + r * ldm);
          la[i] = __ldg(p);
          lb[i] = __ldg(p + 1);
        }
      }
    };
    if (j > 0) load_block(0);
    // b_j is fetched now: its L2 latency must not sit between the last product and the block solve
    double bj[NRHS];
    {
      const int rr = warp * 16 + (lane >> 1);
#pragma unroll
      for (int q = 0; q < NRHS; ++q) bj[q] = (rr < nbj) ? B[(int64_t)q * m + r0 + rr] : 0.0;
    }
    for (int k = 0; k < j; ++k) {
      const Tagged* wk = Wt + (int64_t)k * NRHS * NB;
      if (k + 1 < j) {
        // far from the chain: one backed-off sentinel poll per warp keeps the L2 polling traffic of the ~1000
        // waiting warps small; the block next in the chain (k == j - 1) polls its words directly
        if (lane == 0) (void)ld_tagged(wk, epoch, fault, true);
        __syncwarp();
      }
      double w[NRHS][4];
      poll_tagged<NRHS, 4>(wk + c4, NB, epoch, fault, w);
#pragma unroll
      for (int i = 0; i < 16; ++i)
#pragma unroll
        for (int q = 0; q < NRHS; ++q)
          acc[q][i] += la[i].x * w[q][0] + la[i].y * w[q][1] + lb[i].x * w[q][2] + lb[i].y * w[q][3];
      if (k + 1 < j) load_block(k + 1);  // in registers before w_{k+1} is polled
    }
    // fold across lanes; c = b - sum
#pragma unroll
    for (int q = 0; q < NRHS; ++q) {
      const double tot = fold16(acc[q], lane);
      const int rr = warp * 16 + (lane >> 1);
      if ((lane & 1) == 0) sc[q * NB + rr] = bj[q] - tot;
    }
    __syncthreads();  // c_j and X_j are in shared memory
    double wj[NRHS];
    block_gemv<NRHS, false>(sX, sc, tid, wj);
    if ((tid & 1) == 0) {
#pragma unroll
      for (int q = 0; q < NRHS; ++q) st_tagged(Wt + ((int64_t)j * NRHS + q) * NB + (tid >> 1), wj[q], epoch);
    }
  }

  // ============================================================ backward: L^T x = w
  for (int j = nblk - 1 - blockIdx.x; j >= 0; j -= gridDim.x) {
    const int64_t j0 = (int64_t)j * NB;
    const int nbj = (int)((m - j0) < NB ? (m - j0) : NB);
    __syncthreads();  // previous users of sX / sc / sred are done
    stage_inverse<true>(Xall + (int64_t)j * NB * NB, sX, tid);
    double acc[NRHS][4];
#pragma unroll
    for (int q = 0; q < NRHS; ++q) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0.0;
    double2 la[16], lb[16];
    auto load_block = [&](int k) {
      const double* base = L + j0 + c4;
      const int64_t rw = (int64_t)k * NB + warp * 16;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int64_t r = rw + i;
        la[i] = lb[i] = make_double2(0.0, 0.0);
        if (r < m) {
          const double2* p = reinterpret_cast<const double2*>(base + r * ldm);
          la[i] = __ldg(p);
          lb[i] = __ldg(p + 1);
        }
      }
    };
    if (j + 1 < nblk) load_block(nblk - 1);
    // w_j (from the forward owner of block j, long done except for the very first backward blocks) is fetched
    // now, off the chain
    double wjv[1][1] = {{0.0}};
    if (tid < NRHS * NB)
      poll_tagged<1, 1>(Wt + ((int64_t)j * NRHS + (tid >> 7)) * NB + (tid & (NB - 1)), 0, epoch, fault, wjv);
    for (int k = nblk - 1; k > j; --k) {
      // lane (i = lane & 15, q = lane >> 4) fetches x_k[16 warp + i] of right-hand side q; shuffles hand it round
      const int qi = (NRHS == 2) ? (lane >> 4) : 0;
      const Tagged* xk = Xt + ((int64_t)k * NRHS + qi) * NB + warp * 16 + (lane & 15);
      if (k - 1 > j) {
        if (lane == 0) (void)ld_tagged(xk, epoch, fault, true);
        __syncwarp();
      }
      double mine_v[1][1];
      poll_tagged<1, 1>(xk, 0, epoch, fault, mine_v);
      const double mine = mine_v[0][0];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
#pragma unroll
        for (int q = 0; q < NRHS; ++q) {
          const double xv = __shfl_sync(full, mine, i + 16 * q);
          acc[q][0] += la[i].x * xv;
          acc[q][1] += la[i].y * xv;
          acc[q][2] += lb[i].x * xv;
          acc[q][3] += lb[i].y * xv;
        }
      }
      if (k - 1 > j) load_block(k - 1);
    }
#pragma unroll
    for (int q = 0; q < NRHS; ++q) {
      *reinterpret_cast<double2*>(sred + (warp * 2 + q) * NB + c4) = make_double2(acc[q][0], acc[q][1]);
      *reinterpret_cast<double2*>(sred + (warp * 2 + q) * NB + c4 + 2) = make_double2(acc[q][2], acc[q][3]);
    }
    __syncthreads();
    if (tid < NRHS * NB) {  // c = w_j - sum over the 8 warps (w_j comes from the forward owner of block j)
      const int q = tid >> 7, c = tid & (NB - 1);
      const double wv = wjv[0][0];
      double sum = 0.0;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += sred[(w * 2 + q) * NB + c];
      sc[q * NB + c] = (c < nbj) ? wv - sum : 0.0;
    }
    __syncthreads();
    double xj[NRHS];
    block_gemv<NRHS, true>(sX, sc, tid, xj);
    if ((tid & 1) == 0) {
      const int c = tid >> 1;
#pragma unroll
      for (int q = 0; q < NRHS; ++q) {
        st_tagged(Xt + ((int64_t)j * NRHS + q) * NB + c, xj[q], epoch);
        if (c < nbj) B[(int64_t)q * m + j0 + c] = xj[q];
      }
    }
  }
}

// ------------------------------------------------------------------ distributed factorisation: panel pack / unpack
// Broadcast payload of panel k:  [ inv(L_kk) : 128 x 128 ][ rows k0 .. m-1 of columns k0 .. k0+nb-1 : (m - k0) x nb ],
// contiguous.  TO_BUF: owner, after potf2_inv + TRSM.  !TO_BUF: everyone else, after the broadcast.
template <bool TO_BUF>
__global__ void __launch_bounds__(256)
panel_pack_kernel(double* __restrict__ Mat, int64_t ldm, int64_t k0, int64_t m, int nb, double* __restrict__ Linv,
                  double* __restrict__ buf) {
  const int64_t n_inv = NB * NB;
  const int64_t total = n_inv + (m - k0) * nb;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    double* p;
    if (idx < n_inv) {
      p = Linv + idx;
    } else {
      const int64_t e = idx - n_inv;
      const int64_t r = e / nb, c = e - r * nb;
      p = Mat + (k0 + r) * ldm + k0 + c;
    }
    if (TO_BUF)
      buf[idx] = *p;
    else
      *p = buf[idx];
  }
}

// Packed panel of the two-broadcast distributed factorisation (k_potrf_dist2), all in units of doubles:
//   [0, NB NB)               rows of block k+1        (the SMALL broadcast)
//   [NB NB, 2 NB NB)         the diagonal block L_kk  (dense, zero above the diagonal)
//   [2 NB NB, ... )          rows from block k+2 on
//   then NB NB               inv(L_kk) (its 16 x 16 diagonal sub-blocks)
struct PackedPanel {
  int64_t k0, rows_total, r1, rest, off_linv, large_count, map_rows;
  PackedPanel(int64_t m, int k) {
    k0 = (int64_t)k * NB;
    rows_total = m - k0;
    const int64_t nbk = rows_total < NB ? rows_total : NB;
    const int64_t below = rows_total - nbk;
    r1 = below < NB ? below : NB;
    rest = below - r1;
    off_linv = 2 * (int64_t)NB * NB + rest * NB;
    large_count = (int64_t)NB * NB + rest * NB + (int64_t)NB * NB;
    map_rows = 2 * NB + rest;  // rows of the buffer seen by the update kernels' tensor map
  }
};

// Non-owners: copy a received packed panel into the local M (lower triangle of the diagonal block, all rows below)
// and its inverted diagonal sub-blocks into the Linv workspace.
__global__ void __launch_bounds__(256)
panel_unpack2_kernel(double* __restrict__ Mat, int64_t ldm, int64_t k0, int64_t m, const double* __restrict__ buf,
                     int64_t off_linv, double* __restrict__ Linv) {
  const int64_t rows = m - k0;
  const int64_t total = rows * NB + (int64_t)NB * NB;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    if (idx < rows * NB) {
      const int64_t lr = idx >> 7;
      const int c = (int)(idx & (NB - 1));
      int64_t br = lr;                       // buffer row of local row lr: the first two blocks are swapped
      if (lr < NB) br = lr + NB;
      else if (lr < 2 * NB) br = lr - NB;
      if (lr >= NB || c <= lr) Mat[(k0 + lr) * ldm + k0 + c] = buf[br * NB + c];
    } else {
      const int64_t e = idx - rows * NB;
      Linv[e] = buf[off_linv + e];
    }
  }
}

template <typename K>
int set_smem(K kern, size_t bytes) {
  LPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return LPB_OK;
}

constexpr size_t kPotf2Smem = (size_t)(NB * LDS + NB) * sizeof(double);
constexpr size_t kPotf2InvSmem = (size_t)(NB * LDS + NB + 2 * SB * XDP + 16 * SB * TWP + panel_factor_scratch(XDP)) * sizeof(double);
constexpr size_t kTrsmSmem = (size_t)((NB + TRSM_ROWS) * LDS) * sizeof(double);
constexpr size_t kTrsvSmem = (size_t)(NB * LDS + NB + 2 * NB) * sizeof(double);

int configure_once() {
  static PerDeviceOnce once;
  return once.run([](int) -> int {
  LPB_TRY(set_smem(potf2_kernel, kPotf2Smem));
  LPB_TRY(set_smem(potf2_inv_kernel, kPotf2InvSmem));
  LPB_TRY(set_smem(trsm_kernel, kTrsmSmem));
  LPB_TRY(set_smem(trsm_blocked_kernel, kTrsmBlockedSmem));
  LPB_TRY(set_smem(trsm_dmma_blocked_kernel, kTrsmDmmaSmem));
  LPB_TRY(set_smem(trsv_diag_kernel<false, 1>, kTrsvSmem));
  LPB_TRY(set_smem(trsv_diag_kernel<false, 2>, kTrsvSmem));
  LPB_TRY(set_smem(trsv_diag_kernel<true, 1>, kTrsvSmem));
  LPB_TRY(set_smem(trsv_diag_kernel<true, 2>, kTrsvSmem));
  return LPB_OK;
  });
}

#define LPB_KCHECK(lc)             \
  do {                             \
    (lc).launches++;               \
    LPB_CUDA(cudaGetLastError());  \
  } while (0)

}  // namespace

// Workspace layout (doubles): [ inv(L_jj): nblk x 128 x 128 ][ scratch right-hand sides: 2 ldy ][ flags of the
// flag-based pipelined solve: 2 nblk ints ][ tagged hand-off words of solve_ll_kernel: forward + backward,
// nblk x 2 x 128 x 16 bytes each ].  Zero-filled at allocation: flags and tags start below any epoch.
struct CholWs {
  int64_t nblk, ldy, off_y, off_flags, off_wt, off_xt, total;
  explicit CholWs(int64_t m) {
    nblk = ceil_div(m, NB);
    ldy = round_up(m, 2);
    off_y = nblk * NB * NB;
    off_flags = off_y + 2 * ldy;
    off_wt = off_flags + round_up(nblk, 2);
    off_xt = off_wt + nblk * 2 * NB * 2;
    total = off_xt + nblk * 2 * NB * 2;
  }
};

static int ensure_chol_ws(LaunchCtx& lc, int64_t m) {
  const int64_t need = CholWs(m).total;
  if (lc.chol_ws_cap >= need) return LPB_OK;
  if (lc.chol_ws) cudaFree(lc.chol_ws);
  lc.chol_ws = nullptr;
  lc.chol_ws_cap = 0;
  lc.linv_valid_m = -1;
  void* p = nullptr;
  LPB_CUDA(cudaMalloc(&p, sizeof(double) * (size_t)need));
  LPB_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * (size_t)need, lc.stream));
  lc.chol_ws = static_cast<double*>(p);
  lc.chol_ws_cap = need;
  return LPB_OK;
}

// After a factorisation: complete the 128 x 128 block inverses (all blocks in parallel, ~20 us) for the solves.
static int finish_factor(LaunchCtx& lc, int64_t m, double* Mat, int64_t ldm) {
  if (lc.potf2_impl == 1) {  // no inverted blocks were produced: the solves fall back to plain substitution
    lc.linv_valid_m = -1;
    lc.linv_mat = nullptr;
    return LPB_OK;
  }
  if (!lc.linv_full) {
    static PerDeviceOnce once;
    LPB_TRY(once.run([](int) -> int { return set_smem(linv_complete_kernel, kLinvCompleteSmem); }));
    linv_complete_kernel<<<(unsigned)ceil_div(m, NB), dim3(32, 16), kLinvCompleteSmem, lc.stream>>>(Mat, ldm, m, lc.chol_ws);
    LPB_KCHECK(lc);
    lc.linv_full = true;
  }
  lc.linv_valid_m = m;
  lc.linv_mat = Mat;
  return LPB_OK;
}

static int k_potrf_dist(LaunchCtx& lc, int64_t m, double* Mat, int64_t ldm);
static int k_potrf_dist2(LaunchCtx& lc, int64_t m, double* Mat, int64_t ldm);

// Single-GPU factorisation with look-ahead.  The serial part of a panel is potf2_inv: one CTA, ~40 us, with
// 147 SMs idle.  Here the trailing update of panel k is split: its first block column (the one panel k+1
// lives in) is updated first, then potf2_inv(k+1) runs on a second, high-priority stream while the update of
// the remaining columns keeps 147 SMs busy on the main stream; the TRSM of panel k+1 follows on the main
// stream once both are done.  Same kernels, same order of updates per tile as the sequential loop, so the
// factor is bit-identical to it (option "potrf_lookahead" = 0 selects the sequential loop).
static int k_potrf_lookahead(LaunchCtx& lc, int64_t m, double* Mat, int64_t ldm) {
  if (!lc.side_stream) {
    int lo = 0, hi = 0;
    LPB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LPB_CUDA(cudaStreamCreateWithPriority(&lc.side_stream, cudaStreamNonBlocking, hi));
    for (int e = 0; e < 2; ++e) {
      LPB_CUDA(cudaEventCreateWithFlags(&lc.ev_col[e], cudaEventDisableTiming));
      LPB_CUDA(cudaEventCreateWithFlags(&lc.ev_pan[e], cudaEventDisableTiming));
    }
  }
  const int T = (int)ceil_div(m, NB);
  lc.linv_full = false;
  auto nb_of = [&](int k) { return (int)((m - (int64_t)k * NB) < NB ? (m - (int64_t)k * NB) : NB); };
  auto potf2 = [&](int k, cudaStream_t st) -> int {
    potf2_inv_kernel<<<1, dim3(32, 16), kPotf2InvSmem, st>>>(Mat, ldm, k * NB, nb_of(k), lc.info_dev,
                                                             lc.chol_ws + (int64_t)k * NB * NB, 0);
    LPB_KCHECK(lc);
    return LPB_OK;
  };
  auto trsm = [&](int k) -> int {
    const int64_t rem = m - (int64_t)k * NB - nb_of(k);
    if (rem <= 0) return LPB_OK;
    trsm_dmma_blocked_kernel<<<(unsigned)ceil_div(rem, TBR), 256, kTrsmDmmaSmem, lc.stream>>>(
        Mat, ldm, k * NB, (int)m, lc.chol_ws + (int64_t)k * NB * NB);
    LPB_KCHECK(lc);
    return LPB_OK;
  };
  LPB_TRY(potf2(0, lc.stream));
  LPB_TRY(trsm(0));
  const int saved_cap = lc.update_grid_cap;
  int rc = LPB_OK;
  for (int k = 0; k + 1 < T && rc == LPB_OK; ++k) {
    const int64_t k0 = (int64_t)k * NB;
    cudaEvent_t ec = lc.ev_col[k & 1], ep = lc.ev_pan[k & 1];
    // (1) the block column of panel k+1, then hand it to the side stream
    rc = k_trailing_update_part(lc, m, Mat, ldm, k0, NB, k + 1, 1, 1, 0);
    if (rc != LPB_OK) break;
    LPB_CUDA(cudaEventRecord(ec, lc.stream));
    LPB_CUDA(cudaStreamWaitEvent(lc.side_stream, ec, 0));
    rc = potf2(k + 1, lc.side_stream);
    if (rc != LPB_OK) break;
    LPB_CUDA(cudaEventRecord(ep, lc.side_stream));
    // (2) the rest of update k on all SMs but one
    lc.update_grid_cap = kNumSMs - 1;
    rc = k_trailing_update_part(lc, m, Mat, ldm, k0, NB, k + 2, 0, 1, 0);
    lc.update_grid_cap = saved_cap;
    if (rc != LPB_OK) break;
    // (3) panel k+1 joins the main stream
    LPB_CUDA(cudaStreamWaitEvent(lc.stream, ep, 0));
    rc = trsm(k + 1);
  }
  lc.update_grid_cap = saved_cap;
  LPB_TRY(rc);
  return finish_factor(lc, m, Mat, ldm);
}

int k_potrf(LaunchCtx& lc, int64_t m, double* Mat, int64_t ldm, int syrk_impl) {
  LPB_TRY(configure_once());
  LPB_TRY(ensure_chol_ws(lc, m));
  if (lc.ws_m != m) {  // the workspace layout depends on m: stale words of another layout must not look like tags
    const CholWs ws(m);
    LPB_CUDA(cudaMemsetAsync(lc.chol_ws + ws.off_flags, 0, sizeof(double) * (size_t)(ws.total - ws.off_flags), lc.stream));
    lc.ws_m = m;
  }
  LPB_CUDA(cudaMemsetAsync(lc.info_dev, 0, sizeof(int), lc.stream));
  const bool dmma_update = lc.update_impl == 0 || lc.update_impl == 2 || lc.update_impl == 6;
  if (lc.world > 1 && lc.nccl_comm && lc.potrf_dist && syrk_impl == 0 && lc.trsm_impl == 0 && dmma_update &&
      lc.potf2_impl == 0 && m > NB && !(ldm & 1) && !(reinterpret_cast<uintptr_t>(Mat) & 15))
    return lc.potrf_dist == 2 ? k_potrf_dist2(lc, m, Mat, ldm) : k_potrf_dist(lc, m, Mat, ldm);
  if (lc.potrf_lookahead && syrk_impl == 0 && lc.trsm_impl == 0 && dmma_update && lc.potf2_impl == 0 &&
      !lc.sync_each_launch && m > 2 * NB && !(ldm & 1) && !(reinterpret_cast<uintptr_t>(Mat) & 15))
    return k_potrf_lookahead(lc, m, Mat, ldm);
  const int full_inverse = lc.trsm_impl == 2 ? 1 : 0;  // the solves complete the inverses themselves (finish_factor)
  lc.linv_full = full_inverse != 0;
  for (int64_t k0 = 0; k0 < m; k0 += NB) {
    const int nb = (int)((m - k0) < NB ? (m - k0) : NB);
    const int64_t rem = m - k0 - nb;
    double* linv = lc.chol_ws + (k0 / NB) * NB * NB;
    if (lc.potf2_impl == 1)  // textbook column Cholesky with sqrt and divisions, no inverses (accuracy experiments)
      potf2_kernel<<<1, dim3(32, 16), kPotf2Smem, lc.stream>>>(Mat, ldm, (int)k0, nb, lc.info_dev);
    else
      potf2_inv_kernel<<<1, dim3(32, 16), kPotf2InvSmem, lc.stream>>>(Mat, ldm, (int)k0, nb, lc.info_dev, linv,
                                                                      full_inverse);
    LPB_KCHECK(lc);
    if (lc.sync_each_launch) LPB_CUDA(cudaStreamSynchronize(lc.stream));
    if (rem > 0) {
      if (syrk_impl == 1 || lc.trsm_impl == 1) {  // column-by-column substitution (reference path)
        trsm_kernel<<<(unsigned)ceil_div(rem, TRSM_ROWS), dim3(32, 8), kTrsmSmem, lc.stream>>>(Mat, ldm, (int)k0, nb,
                                                                                               (int)m);
        LPB_KCHECK(lc);
      } else if (lc.trsm_impl == 2) {             // GEMM with the full 128 x 128 inverse (not backward stable)
        LPB_TRY(k_trsm_dmma(lc, m, Mat, ldm, k0, linv));
      } else if (lc.trsm_impl == 3 || (ldm & 1) || (reinterpret_cast<uintptr_t>(Mat) & 15)) {
        // blocked substitution in plain DFMA: reference for the default, and the path for unaligned matrices
        trsm_blocked_kernel<<<(unsigned)ceil_div(rem, TBR), 256, kTrsmBlockedSmem, lc.stream>>>(Mat, ldm, (int)k0,
                                                                                               (int)m, linv);
        LPB_KCHECK(lc);
      } else {                                    // rem > 0 implies nb == NB
        trsm_dmma_blocked_kernel<<<(unsigned)ceil_div(rem, TBR), 256, kTrsmDmmaSmem, lc.stream>>>(Mat, ldm, (int)k0,
                                                                                                 (int)m, linv);
        LPB_KCHECK(lc);
      }
      if (lc.sync_each_launch) LPB_CUDA(cudaStreamSynchronize(lc.stream));
      if (syrk_impl == 1 || lc.update_impl == 1)
        LPB_TRY(k_trailing_update_simple(lc, m, Mat, ldm, k0, nb));
      else
        LPB_TRY(k_trailing_update_dmma(lc, m, Mat, ldm, k0, nb));
      if (lc.sync_each_launch) LPB_CUDA(cudaStreamSynchronize(lc.stream));
    }
  }
  return finish_factor(lc, m, Mat, ldm);
}

// Distributed factorisation for column-sharded contexts (every rank holds the same all-reduced M).
// Block column k belongs to rank k mod G.  Its owner factors the diagonal block and solves the panel
// (potf2_inv + TRSM, as in k_potrf), then BROADCASTS the finished panel together with inv(L_kk) over
// NVLink (one ncclBroadcast per panel); every rank stores it into its own copy of M, so at the end all
// ranks hold the whole factor and the triangular solves stay local.  The trailing update -- the m^3/3
// flop -- is sharded: each rank updates only the block columns it owns.  The rank that owns panel k+1
// updates that column first and defers the rest of its update until its panel is on the wire, so its
// potf2 / TRSM chain overlaps the other ranks' updates (look-ahead across ranks, one stream per rank).
// Bit-identical factors on all ranks by construction (every entry of L is computed by exactly one rank).
static int ensure_dist_streams(LaunchCtx& lc) {
  if (!lc.side_stream) {
    int lo = 0, hi = 0;
    LPB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    LPB_CUDA(cudaStreamCreateWithPriority(&lc.side_stream, cudaStreamNonBlocking, hi));
    for (int e = 0; e < 2; ++e) {
      LPB_CUDA(cudaEventCreateWithFlags(&lc.ev_col[e], cudaEventDisableTiming));
      LPB_CUDA(cudaEventCreateWithFlags(&lc.ev_pan[e], cudaEventDisableTiming));
    }
  }
  for (int e = 0; e < 4; ++e)
    if (!lc.ev_dist[e]) LPB_CUDA(cudaEventCreateWithFlags(&lc.ev_dist[e], cudaEventDisableTiming));
  return LPB_OK;
}

// ------------------------------------------------------------------ peer-memory hand-off of the first 128 panel rows
// The chain of the distributed factorisation runs  owner(k): potf2 -> TRSM -> [128 rows of block k + 1 to everybody]
// -> owner(k + 1): one-tile update -> potf2 ...  Through ncclBroadcast the bracket costs ~30 us at 8 ranks (launch and
// protocol latency for 128 KB); here the owner's stream runs peer_push_kernel, which writes the rows into every
// peer's panel slot over NVLink (16-byte stores to cudaIpc-mapped memory), and the last CTA to finish raises a flag in
// each peer's memory; the peers' streams run wait_flag_kernel.  No receive-side kernel, no staging buffer.
// Flow control is by data dependency: the ring holds 2 x world slots, and the push of panel k + 2 world cannot start
// before every rank has pushed a panel of its own in between -- which, in stream order, comes after all its reads of
// panel k.  Consecutive factorisations are separated by a collective (k_potrf_dist2 starts with a 1-word all-reduce).
constexpr int kPushCtasPerPeer = 4;
constexpr int kRingHeaderDoubles = 128;  // 1 KB: flags (one u64 per slot) + the push kernel's CTA counter
struct PeerPush {
  double* dst[LaunchCtx::kMaxPeers - 1];
  unsigned long long* flag[LaunchCtx::kMaxPeers - 1];
  int n;
};
__global__ void __launch_bounds__(256)
peer_push_kernel(const double* __restrict__ src, int64_t count2, PeerPush pp, unsigned long long value,
                 unsigned* __restrict__ counter) {
  const int peer = blockIdx.x / kPushCtasPerPeer, part = blockIdx.x % kPushCtasPerPeer;
  const double2* s = reinterpret_cast<const double2*>(src);
  double2* d = reinterpret_cast<double2*>(pp.dst[peer]);
  for (int64_t e = part * 256 + threadIdx.x; e < count2; e += kPushCtasPerPeer * 256) d[e] = s[e];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {  // every CTA's stores are fenced and counted: publish
      *counter = 0;               // for the next launch (stream-ordered behind this one)
      __threadfence_system();
      for (int g = 0; g < pp.n; ++g)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(pp.flag[g]), "l"(value) : "memory");
    }
  }
}
// One thread polls this rank's flag of the slot; a lost hand-off raises the context's fault word after ~10 s instead
// of hanging (the host then reports LPB_ERR_CUDA at its next scalar fetch), and a fault raised elsewhere ends the wait.
__global__ void wait_flag_kernel(const unsigned long long* __restrict__ flag, unsigned long long value,
                                 unsigned long long* __restrict__ fault) {
  if (threadIdx.x != 0) return;
  unsigned long long t0, t1, v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    if (v >= value) return;
    __nanosleep(100);
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 10000000000ull) {
      if (fault) atomicExch(fault, 1ull);
      return;
    }
    if (fault && *reinterpret_cast<volatile unsigned long long*>(fault) != 0ull) return;
  }
}

static int64_t dist_slot_doubles(int64_t m) { return 3 * (int64_t)NB * NB + round_up(m, NB) * NB; }

// The ring and its mappings belong to the PROCESS (one process per GPU), like the NCCL communicator: cudaMalloc +
// cudaIpcOpenMemHandle + the handle exchange cost ~0.4 s, far too much for a context that lives for one solve
// (InteriorPoint.solve on a sharded problem creates one per call).  A context attaches to the ring while it lives;
// a second live context of the same process finds it busy and keeps the ncclBroadcast path.  The epoch of the flag
// values is process-wide too: the flags in the ring outlive contexts.
struct ProcessRing {
  std::mutex mu;
  double* base = nullptr;
  int64_t slot_doubles = 0;
  int slots = 0, world = 0, rank = -1;
  double* peer[LaunchCtx::kMaxPeers] = {};
  bool mapped = false;
  const LaunchCtx* owner = nullptr;
  uint32_t epoch = 0;
};
static ProcessRing g_ring;

static void ring_free_locked() {
  for (int g = 0; g < LaunchCtx::kMaxPeers; ++g) {
    if (g_ring.peer[g] && g_ring.peer[g] != g_ring.base) cudaIpcCloseMemHandle(g_ring.peer[g]);
    g_ring.peer[g] = nullptr;
  }
  if (g_ring.base) cudaFree(g_ring.base);
  g_ring.base = nullptr;
  g_ring.slots = 0;
  g_ring.slot_doubles = 0;
  g_ring.mapped = false;
}

static void ring_attach_locked(LaunchCtx& lc) {
  lc.ring_base = g_ring.base;
  lc.ring_slots = g_ring.slots;
  lc.ring_slot_doubles = g_ring.slot_doubles;
  for (int g = 0; g < LaunchCtx::kMaxPeers; ++g) lc.peer_base[g] = g_ring.peer[g];
  lc.peer_mapped = g_ring.mapped;
  lc.peer_ready = g_ring.mapped;
  g_ring.owner = &lc;
}

void k_peer_release(LaunchCtx& lc) {
  std::lock_guard<std::mutex> lk(g_ring.mu);
  if (g_ring.owner == &lc) g_ring.owner = nullptr;
  lc.ring_base = nullptr;
  for (int g = 0; g < LaunchCtx::kMaxPeers; ++g) lc.peer_base[g] = nullptr;
  lc.ring_slots = 0;
  lc.ring_slot_doubles = 0;
  lc.peer_mapped = false;
  lc.peer_ready = false;
}

void k_peer_free_process() {
  std::lock_guard<std::mutex> lk(g_ring.mu);
  if (!g_ring.owner) ring_free_locked();
}

// *state_out: 0 = attached to the ring this process has already mapped (nothing to exchange); 1 = a new ring was
// allocated, exchange the handles and call k_peer_import; 2 = another live context of this process holds the ring:
// no peer hand-off for this one.  Ranks with the same history of contexts get the same answer.
int k_peer_export(LaunchCtx& lc, int64_t m, unsigned char* handle_out, int* state_out) {
  if (lc.world < 2 || lc.world > LaunchCtx::kMaxPeers || !handle_out || !state_out) {
    set_last_error("peer_export: needs a sharded context of 2..%d ranks", LaunchCtx::kMaxPeers);
    return LPB_ERR_BAD_ARGUMENT;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::lock_guard<std::mutex> lk(g_ring.mu);
  if (g_ring.owner && g_ring.owner != &lc) {
    *state_out = 2;
    return LPB_OK;
  }
  const int64_t per = dist_slot_doubles(m);
  if (g_ring.base && g_ring.mapped && g_ring.world == lc.world && g_ring.rank == lc.rank && g_ring.slot_doubles >= per) {
    ring_attach_locked(lc);
    *state_out = 0;
    return LPB_OK;
  }
  ring_free_locked();
  const int slots = 2 * lc.world;
  const size_t bytes = sizeof(double) * (size_t)(kRingHeaderDoubles + slots * per);
  void* p = nullptr;
  LPB_CUDA(cudaMalloc(&p, bytes));
  LPB_CUDA(cudaMemsetAsync(p, 0, bytes, lc.stream));
  LPB_CUDA(cudaStreamSynchronize(lc.stream));  // the peers may write as soon as they have the handle
  g_ring.base = static_cast<double*>(p);
  g_ring.slots = slots;
  g_ring.slot_doubles = per;
  g_ring.world = lc.world;
  g_ring.rank = lc.rank;
  g_ring.owner = &lc;
  cudaIpcMemHandle_t h;
  LPB_CUDA(cudaIpcGetMemHandle(&h, p));
  memcpy(handle_out, &h, sizeof(h));
  *state_out = 1;
  return LPB_OK;
}

int k_peer_import(LaunchCtx& lc, const unsigned char* handles, int world) {
  std::lock_guard<std::mutex> lk(g_ring.mu);
  if (!g_ring.base || g_ring.owner != &lc || world != lc.world || !handles) {
    set_last_error("peer_import: call peer_export first (world %d)", lc.world);
    return LPB_ERR_BAD_ARGUMENT;
  }
  for (int g = 0; g < world; ++g) {
    if (g == lc.rank) {
      g_ring.peer[g] = g_ring.base;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + 64 * (size_t)g, sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_last_error("peer_import: cudaIpcOpenMemHandle of rank %d -> %s", g, cudaGetErrorString(e));
      ring_free_locked();
      g_ring.owner = nullptr;
      return LPB_ERR_CUDA;
    }
    g_ring.peer[g] = static_cast<double*>(p);
  }
  g_ring.mapped = true;
  ring_attach_locked(lc);
  return LPB_OK;
}

// every rank decided to drop the ring (somebody failed to export or import)
void k_peer_abandon(LaunchCtx& lc) {
  {
    std::lock_guard<std::mutex> lk(g_ring.mu);
    if (g_ring.owner == &lc || !g_ring.owner) {
      g_ring.owner = nullptr;
      ring_free_locked();
    }
  }
  k_peer_release(lc);
}

int k_potrf_dist_reserve(LaunchCtx& lc, int64_t m) {
  const int64_t slot = 3 * (int64_t)NB * NB + round_up(m, NB) * NB;
  if (lc.panel_slot_cap < slot) {
    for (int i = 0; i < 2; ++i) {
      if (lc.panel_slot[i]) cudaFree(lc.panel_slot[i]);
      lc.panel_slot[i] = nullptr;
    }
    lc.panel_slot_cap = 0;
    for (int i = 0; i < 2; ++i) {
      void* p = nullptr;
      LPB_CUDA(cudaMalloc(&p, sizeof(double) * (size_t)slot));
      LPB_CUDA(cudaMemsetAsync(p, 0, sizeof(double) * (size_t)slot, lc.stream));
      lc.panel_slot[i] = static_cast<double*>(p);
    }
    lc.panel_slot_cap = slot;
  }
  LPB_TRY(ensure_dist_streams(lc));
  const int64_t need = (int64_t)NB * NB + m * NB;
  if (lc.panel_buf_cap >= need) return LPB_OK;
  if (lc.panel_buf) cudaFree(lc.panel_buf);
  lc.panel_buf = nullptr;
  lc.panel_buf_cap = 0;
  void* p = nullptr;
  LPB_CUDA(cudaMalloc(&p, sizeof(double) * (size_t)need));
  lc.panel_buf = static_cast<double*>(p);
  lc.panel_buf_cap = need;
  return LPB_OK;
}

static int k_potrf_dist(LaunchCtx& lc, int64_t m, double* Mat, int64_t ldm) {
  const int G = lc.world, me = lc.rank;
  ncclComm_t comm = static_cast<ncclComm_t>(lc.nccl_comm);
  LPB_TRY(k_potrf_dist_reserve(lc, m));  // no-op for contexts made by lpb_create_sharded
  const int T = (int)ceil_div(m, NB);
  lc.linv_full = false;
  struct Ops {
    LaunchCtx& lc;
    ncclComm_t comm;
    int64_t m, ldm;
    double* Mat;
    int G, me;
    int64_t k0(int k) const { return (int64_t)k * NB; }
    int nb(int k) const { return (int)((m - k0(k)) < NB ? (m - k0(k)) : NB); }
    double* linv(int k) const { return lc.chol_ws + (int64_t)k * NB * NB; }
    int64_t count(int k) const { return (int64_t)NB * NB + (m - k0(k)) * nb(k); }
    unsigned pack_grid(int k) const { return (unsigned)std::min<int64_t>(ceil_div(count(k), 256 * 4), kNumSMs * 4); }
    int factor_panel(int k) {
      const int64_t rem = m - k0(k) - nb(k);
      potf2_inv_kernel<<<1, dim3(32, 16), kPotf2InvSmem, lc.stream>>>(Mat, ldm, (int)k0(k), nb(k), lc.info_dev, linv(k), 0);
      LPB_KCHECK(lc);
      if (rem > 0) {
        trsm_dmma_blocked_kernel<<<(unsigned)ceil_div(rem, TBR), 256, kTrsmDmmaSmem, lc.stream>>>(Mat, ldm, (int)k0(k),
                                                                                                 (int)m, linv(k));
        LPB_KCHECK(lc);
      }
      panel_pack_kernel<true><<<pack_grid(k), 256, 0, lc.stream>>>(Mat, ldm, k0(k), m, nb(k), linv(k), lc.panel_buf);
      LPB_KCHECK(lc);
      return LPB_OK;
    }
    int broadcast(int k, int owner) {
      const ncclResult_t r =
          ncclBroadcast(lc.panel_buf, lc.panel_buf, (size_t)count(k), ncclDouble, owner, comm, lc.stream);
      if (r != ncclSuccess) {
        set_last_error("potrf_dist: ncclBroadcast of panel %d -> %s", k, ncclGetErrorString(r));
        return LPB_ERR_NCCL;
      }
      return LPB_OK;
    }
    int store_panel(int k) {
      panel_pack_kernel<false><<<pack_grid(k), 256, 0, lc.stream>>>(Mat, ldm, k0(k), m, nb(k), linv(k), lc.panel_buf);
      LPB_KCHECK(lc);
      return LPB_OK;
    }
    int update_column(int p, int col) { return k_trailing_update_part(lc, m, Mat, ldm, k0(p), nb(p), col, 1, 1, 0); }
    int update_owned(int p, int tile0) { return k_trailing_update_part(lc, m, Mat, ldm, k0(p), nb(p), tile0, 0, G, me); }
  } ops{lc, comm, m, ldm, Mat, G, me};
  LPB_TRY(potrf_dist_schedule(T, G, me, ops));
  return finish_factor(lc, m, Mat, ldm);
}

// Schedule v2 of the distributed factorisation (dist_schedule.hpp::potrf_dist_schedule2): the panel is produced
// straight into a packed buffer (slot k & 1), travels as two broadcasts (first the 128 rows the next owner needs
// for its diagonal tile, then the rest), the updates read it from the buffer through a TMA map, and potf2 of the
// next panel runs on the side stream while the rest of the panel is still on the wire.  Same kernels, same order
// of updates per tile: the factor is bit-identical to the single-GPU one and to schedule 1.
static int k_potrf_dist2(LaunchCtx& lc, int64_t m, double* Mat, int64_t ldm) {
  const int G = lc.world, me = lc.rank;
  ncclComm_t comm = static_cast<ncclComm_t>(lc.nccl_comm);
  LPB_TRY(k_potrf_dist_reserve(lc, m));
  const int T = (int)ceil_div(m, NB);
  lc.linv_full = false;
  // peer-memory hand-off of the first 128 rows of every panel (see peer_push_kernel): only if every rank mapped every
  // other rank's ring and the ring's slots hold a panel of this m
  const bool peer = lc.peer_ready && lc.ring_slots >= 2 * G && lc.ring_slot_doubles >= dist_slot_doubles(m);
  if (peer) {
    // no rank may push into a slot a slower rank still reads from the PREVIOUS factorisation: a 1-word all-reduce is
    // stream-ordered behind every rank's earlier kernels (the normal iteration has collectives in between anyway)
    {
      std::lock_guard<std::mutex> lk(g_ring.mu);
      lc.dist_epoch = ++g_ring.epoch;
    }
    unsigned long long* scratch = reinterpret_cast<unsigned long long*>(lc.ring_base) + 64;  // header word, unused otherwise
    if (ncclAllReduce(scratch, scratch, 1, ncclUint64, ncclMax, comm, lc.stream) != ncclSuccess) {
      set_last_error("potrf_dist2: barrier all-reduce failed");
      return LPB_ERR_NCCL;
    }
  }
  struct Ops {
    LaunchCtx& lc;
    ncclComm_t comm;
    int64_t m, ldm;
    double* Mat;
    int G, me;
    bool peer;
    cudaStream_t st(int side) const { return side ? lc.side_stream : lc.stream; }
    int nb(int k) const { return (int)((m - (int64_t)k * NB) < NB ? (m - (int64_t)k * NB) : NB); }
    double* linv(int k) const { return lc.chol_ws + (int64_t)k * NB * NB; }
    int64_t slot_off(int k) const { return kRingHeaderDoubles + (int64_t)(k % lc.ring_slots) * lc.ring_slot_doubles; }
    double* slot(int k) const { return peer ? lc.ring_base + slot_off(k) : lc.panel_slot[k & 1]; }
    int potf2(int k, int side) {
      const PackedPanel pp(m, k);
      potf2_inv_kernel<<<1, dim3(32, 16), kPotf2InvSmem, st(side)>>>(Mat, ldm, k * NB, nb(k), lc.info_dev, linv(k), 0,
                                                                     slot(k) + (int64_t)NB * NB, slot(k) + pp.off_linv);
      LPB_KCHECK(lc);
      return LPB_OK;
    }
    int trsm(int k) {
      const int64_t rem = m - (int64_t)k * NB - nb(k);
      if (rem <= 0) return LPB_OK;
      trsm_dmma_blocked_kernel<<<(unsigned)ceil_div(rem, TBR), 256, kTrsmDmmaSmem, lc.stream>>>(Mat, ldm, k * NB, (int)m,
                                                                                               linv(k), slot(k));
      LPB_KCHECK(lc);
      return LPB_OK;
    }
    int bcast(double* p, int64_t count, int owner, int k, const char* what) {
      if (count <= 0) return LPB_OK;
      const ncclResult_t r = ncclBroadcast(p, p, (size_t)count, ncclDouble, owner, comm, lc.stream);
      if (r != ncclSuccess) {
        set_last_error("potrf_dist2: ncclBroadcast (%s) of panel %d -> %s", what, k, ncclGetErrorString(r));
        return LPB_ERR_NCCL;
      }
      return LPB_OK;
    }
    int bcast_small(int k, int owner) {
      const int64_t count = PackedPanel(m, k).r1 * NB;
      if (!peer) return bcast(slot(k), count, owner, k, "block k+1");
      if (count <= 0) return LPB_OK;
      const unsigned long long value = ((unsigned long long)lc.dist_epoch << 32) | (unsigned)(k + 1);
      const int fl = k % lc.ring_slots;
      if (owner == me) {
        PeerPush pp;
        pp.n = 0;
        for (int g = 0; g < G; ++g) {
          if (g == me) continue;
          pp.dst[pp.n] = lc.peer_base[g] + slot_off(k);
          pp.flag[pp.n] = reinterpret_cast<unsigned long long*>(lc.peer_base[g]) + fl;
          ++pp.n;
        }
        unsigned* counter = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned long long*>(lc.ring_base) + 96);
        peer_push_kernel<<<kPushCtasPerPeer * pp.n, 256, 0, lc.stream>>>(slot(k), count / 2, pp, value, counter);
      } else {
        wait_flag_kernel<<<1, 32, 0, lc.stream>>>(reinterpret_cast<unsigned long long*>(lc.ring_base) + fl, value,
                                                  lc.fault_dev);
      }
      LPB_KCHECK(lc);
      return LPB_OK;
    }
    int bcast_large(int k, int owner) {
      return bcast(slot(k) + (int64_t)NB * NB, PackedPanel(m, k).large_count, owner, k, "rest of the panel");
    }
    int record(int ev, int side) {
      LPB_CUDA(cudaEventRecord(lc.ev_dist[ev], st(side)));
      return LPB_OK;
    }
    int wait(int ev, int side) {
      LPB_CUDA(cudaStreamWaitEvent(st(side), lc.ev_dist[ev], 0));
      return LPB_OK;
    }
    int update(int p, int tile0, int single_col, int mod, int rem, int side) {
      lc.launch_on_side = side != 0;
      const int rc = k_trailing_update_part(lc, m, Mat, ldm, (int64_t)p * NB, nb(p), tile0, single_col, mod, rem, slot(p),
                                            PackedPanel(m, p).map_rows);
      lc.launch_on_side = false;
      return rc;
    }
    int update_diag(int p, int col, int side) { return update(p, col, 2, 1, 0, side); }
    int update_col_below(int p, int col) { return update(p, col, 3, 1, 0, 0); }
    int update_owned(int p, int tile0) { return update(p, tile0, 0, G, me, 0); }
    int unpack(int k) {
      const PackedPanel pp(m, k);
      const int64_t total = pp.rows_total * NB + (int64_t)NB * NB;
      const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(total, 256 * 4), kNumSMs * 4);
      panel_unpack2_kernel<<<grid, 256, 0, lc.stream>>>(Mat, ldm, pp.k0, m, slot(k), pp.off_linv, linv(k));
      LPB_KCHECK(lc);
      return LPB_OK;
    }
  } ops{lc, comm, m, ldm, Mat, G, me, peer};
  const int rc = potrf_dist_schedule2(T, G, me, ops);
  lc.launch_on_side = false;
  LPB_TRY(rc);
  return finish_factor(lc, m, Mat, ldm);
}

// Fast path: one launch per 128-block step, using the stored inv(L_kk) blocks of the last k_potrf.
template <int NRHS>
static int potrs_fused(LaunchCtx& lc, int64_t m, const double* L, int64_t ldm, double* B) {
  const int nblk = (int)ceil_div(m, NB);
  double* Y = lc.chol_ws + CholWs(m).off_y;          // scratch m x NRHS (column-major, ld m)
  for (int kb = 0; kb < nblk; ++kb) {                // forward: B -> Y
    const int64_t k0 = (int64_t)kb * NB;
    const int nb = (int)((m - k0) < NB ? (m - k0) : NB);
    const unsigned grid = 1u + (unsigned)ceil_div(m - k0 - nb, NB);
    solve_fwd_step_kernel<NRHS><<<grid, 256, 0, lc.stream>>>(L, ldm, lc.chol_ws + (int64_t)kb * NB * NB, (int)k0, nb, B,
                                                             Y, m);
    LPB_KCHECK(lc);
  }
  for (int kb = nblk - 1; kb >= 0; --kb) {           // backward: Y -> B
    const int64_t k0 = (int64_t)kb * NB;
    const int nb = (int)((m - k0) < NB ? (m - k0) : NB);
    solve_bwd_step_kernel<NRHS><<<1u + (unsigned)kb, 256, 0, lc.stream>>>(L, ldm, lc.chol_ws + (int64_t)kb * NB * NB,
                                                                          (int)k0, nb, Y, B, m);
    LPB_KCHECK(lc);
  }
  return LPB_OK;
}

// One cooperative launch for the whole solve (see solve_pipelined_kernel).
template <int NRHS>
static int potrs_pipelined(LaunchCtx& lc, int64_t m, const double* L, int64_t ldm, double* B) {
  static PerDeviceOnce once;
  static int max_ctas_dev[kMaxDevices] = {};
  auto kern = solve_pipelined_kernel<NRHS>;
  LPB_TRY(once.run([&](int dev) -> int {
    LPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveSmem));
    int sms = 0, per_sm = 0;
    LPB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    LPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSolveThreads, kSolveSmem));
    if (per_sm < 1 || dev < 0 || dev >= kMaxDevices) {
      set_last_error("solve_pipelined_kernel does not fit on this device");
      return LPB_ERR_CUDA;
    }
    max_ctas_dev[dev] = sms * per_sm;
    return LPB_OK;
  }));
  int cur_dev = 0;
  LPB_CUDA(cudaGetDevice(&cur_dev));
  const int max_ctas = max_ctas_dev[cur_dev];
  const CholWs ws(m);
  int nblk = (int)ws.nblk;
  double* Y = lc.chol_ws + ws.off_y;
  int64_t ldy = ws.ldy;
  int* flags_f = reinterpret_cast<int*>(lc.chol_ws + ws.off_flags);
  int* flags_b = flags_f + nblk;
  int epoch = ++lc.solve_epoch;
  const double* linv = lc.chol_ws;
  unsigned long long* fault = lc.fault_dev;
  void* args[] = {(void*)&L, (void*)&ldm, (void*)&linv, (void*)&m, (void*)&nblk, (void*)&B,
                  (void*)&Y, (void*)&ldy, (void*)&flags_f, (void*)&flags_b, (void*)&epoch, (void*)&fault};
  int grid = nblk < max_ctas ? nblk : max_ctas;
  if (lc.solve_grid_cap > 0 && grid > lc.solve_grid_cap) grid = lc.solve_grid_cap;
  LPB_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kSolveThreads), args, kSolveSmem, lc.stream));
  lc.launches++;
  return LPB_OK;
}

// solve_ll_kernel: one cooperative launch, tagged hand-off, full block inverses (the default).
template <int NRHS>
static int potrs_ll(LaunchCtx& lc, int64_t m, const double* L, int64_t ldm, double* B) {
  static PerDeviceOnce once;
  static int max_ctas_dev[kMaxDevices] = {};
  auto kern = solve_ll_kernel<NRHS>;
  LPB_TRY(once.run([&](int dev) -> int {
    LPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSolveLLSmem));
    int sms = 0, per_sm = 0;
    LPB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    LPB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSolveThreads, kSolveLLSmem));
    if (per_sm < 1 || dev < 0 || dev >= kMaxDevices) {
      set_last_error("solve_ll_kernel does not fit on this device");
      return LPB_ERR_CUDA;
    }
    max_ctas_dev[dev] = sms * per_sm;
    return LPB_OK;
  }));
  int cur_dev = 0;
  LPB_CUDA(cudaGetDevice(&cur_dev));
  const int max_ctas = max_ctas_dev[cur_dev];
  const CholWs ws(m);
  int nblk = (int)ws.nblk;
  Tagged* Wt = reinterpret_cast<Tagged*>(lc.chol_ws + ws.off_wt);
  Tagged* Xt = reinterpret_cast<Tagged*>(lc.chol_ws + ws.off_xt);
  uint32_t epoch = (uint32_t)(++lc.solve_epoch);
  const double* xall = lc.chol_ws;
  unsigned long long* fault = lc.fault_dev;
  void* args[] = {(void*)&L, (void*)&ldm, (void*)&xall, (void*)&m, (void*)&nblk, (void*)&B,
                  (void*)&Wt, (void*)&Xt, (void*)&epoch, (void*)&fault};
  int grid = nblk < max_ctas ? nblk : max_ctas;
  if (lc.solve_grid_cap > 0 && grid > lc.solve_grid_cap) grid = lc.solve_grid_cap;
  LPB_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kSolveThreads), args, kSolveLLSmem, lc.stream));
  lc.launches++;
  return LPB_OK;
}

template <int NRHS>
static int potrs_impl(LaunchCtx& lc, int64_t m, const double* L, int64_t ldm, double* B) {
  const int nblk = (int)ceil_div(m, NB);
  // forward: L w = b
  for (int kb = 0; kb < nblk; ++kb) {
    const int64_t k0 = (int64_t)kb * NB;
    const int nb = (int)((m - k0) < NB ? (m - k0) : NB);
    trsv_diag_kernel<false, NRHS><<<1, 512, kTrsvSmem, lc.stream>>>(L, ldm, (int)k0, nb, B, m);
    LPB_KCHECK(lc);
    const int64_t rem = m - k0 - nb;
    if (rem > 0) {
      trsv_update_fwd_kernel<NRHS><<<(unsigned)ceil_div(rem, 64), 256, 0, lc.stream>>>(L, ldm, (int)k0, nb, B, m);
      LPB_KCHECK(lc);
    }
  }
  // backward: L^T x = w
  for (int kb = nblk - 1; kb >= 0; --kb) {
    const int64_t k0 = (int64_t)kb * NB;
    const int nb = (int)((m - k0) < NB ? (m - k0) : NB);
    trsv_diag_kernel<true, NRHS><<<1, 512, kTrsvSmem, lc.stream>>>(L, ldm, (int)k0, nb, B, m);
    LPB_KCHECK(lc);
    if (k0 > 0) {
      trsv_update_bwd_kernel<NRHS><<<(unsigned)ceil_div(k0, 128), dim3(128, 4), 0, lc.stream>>>(L, ldm, (int)k0, nb,
                                                                                               B, m);
      LPB_KCHECK(lc);
    }
  }
  return LPB_OK;
}

int k_potrs(LaunchCtx& lc, int64_t m, const double* L, int64_t ldm, double* B, int nrhs, bool use_linv) {
  LPB_TRY(configure_once());
  const bool aligned = !(ldm & 1) && !(reinterpret_cast<uintptr_t>(L) & 15);  // double2 loads of L
  // solve_impl: 0 = solve_ll_kernel (default), 1 = one launch per block step with the full inverses, 2 = plain
  // substitution (no stored inverses at all), 3 = the flag-based pipelined kernel with blocked substitution over
  // the 16 x 16 inverted sub-blocks (the round-1 default; kept as the accuracy / timing comparison).
  if (use_linv && aligned && lc.solve_impl != 2 && lc.linv_valid_m == m && lc.linv_mat == L && lc.chol_ws &&
      lc.linv_full && lc.fault_dev) {
    if (lc.solve_impl == 1) {
      if (nrhs == 1) return potrs_fused<1>(lc, m, L, ldm, B);
      if (nrhs == 2) return potrs_fused<2>(lc, m, L, ldm, B);
    } else if (lc.solve_impl == 3) {
      if (nrhs == 1) return potrs_pipelined<1>(lc, m, L, ldm, B);
      if (nrhs == 2) return potrs_pipelined<2>(lc, m, L, ldm, B);
    } else {
      if (nrhs == 1) return potrs_ll<1>(lc, m, L, ldm, B);
      if (nrhs == 2) return potrs_ll<2>(lc, m, L, ldm, B);
    }
  } else {
    if (nrhs == 1) return potrs_impl<1>(lc, m, L, ldm, B);
    if (nrhs == 2) return potrs_impl<2>(lc, m, L, ldm, B);
  }
  set_last_error("potrs: nrhs must be 1 or 2");
  return LPB_ERR_BAD_ARGUMENT;
}

}  // namespace lpb
