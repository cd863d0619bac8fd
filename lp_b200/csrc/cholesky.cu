// cholesky.cu -- K2: blocked right-looking lower Cholesky of M (row-major, in place) and
// K3: the forward/back substitutions, 1 or 2 right-hand sides at once.
//
// Replaces /root/reference/src/solvers/interior_point/newton_equations.rs:129-132 (`M.cholesky()`,
// lower L) / :87-90 (LAPACK potrf) and :151-169 (`factor.solvec_into`) / :98-104 (potrs).
// A non-positive or non-finite pivot sets the device `info` flag -> LPB_ERR_NUMERICAL_PROBLEM
// (newton_equations.rs:63); there is no fallback chain.
//
// Per 128-wide panel:  potf2 (one CTA, block in shared memory)  ->  TRSM of the rows below
// (warp-cooperative substitution, 64 rows per CTA)  ->  trailing update C -= P P^T on the DMMA
// path (dmma_gemm.cu).  The solves walk the same 128-blocks: a one-CTA triangular solve of the
// diagonal block, then a coalesced rank-128 update of the remaining right-hand side.
#include "kernels.hpp"

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

template <typename K>
int set_smem(K kern, size_t bytes) {
  LPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return LPB_OK;
}

constexpr size_t kPotf2Smem = (size_t)(NB * LDS + NB) * sizeof(double);
constexpr size_t kTrsmSmem = (size_t)((NB + TRSM_ROWS) * LDS) * sizeof(double);
constexpr size_t kTrsvSmem = (size_t)(NB * LDS + NB + 2 * NB) * sizeof(double);

int configure_once() {
  static bool done = false;
  if (done) return LPB_OK;
  LPB_TRY(set_smem(potf2_kernel, kPotf2Smem));
  LPB_TRY(set_smem(trsm_kernel, kTrsmSmem));
  LPB_TRY(set_smem(trsv_diag_kernel<false, 1>, kTrsvSmem));
  LPB_TRY(set_smem(trsv_diag_kernel<false, 2>, kTrsvSmem));
  LPB_TRY(set_smem(trsv_diag_kernel<true, 1>, kTrsvSmem));
  LPB_TRY(set_smem(trsv_diag_kernel<true, 2>, kTrsvSmem));
  done = true;
  return LPB_OK;
}

#define LPB_KCHECK(lc)             \
  do {                             \
    (lc).launches++;               \
    LPB_CUDA(cudaGetLastError());  \
  } while (0)

}  // namespace

int k_potrf(LaunchCtx& lc, int64_t m, double* Mat, int64_t ldm, int syrk_impl) {
  LPB_TRY(configure_once());
  LPB_CUDA(cudaMemsetAsync(lc.info_dev, 0, sizeof(int), lc.stream));
  for (int64_t k0 = 0; k0 < m; k0 += NB) {
    const int nb = (int)((m - k0) < NB ? (m - k0) : NB);
    potf2_kernel<<<1, dim3(32, 16), kPotf2Smem, lc.stream>>>(Mat, ldm, (int)k0, nb, lc.info_dev);
    LPB_KCHECK(lc);
    const int64_t rem = m - k0 - nb;
    if (rem > 0) {
      trsm_kernel<<<(unsigned)ceil_div(rem, TRSM_ROWS), dim3(32, 8), kTrsmSmem, lc.stream>>>(Mat, ldm, (int)k0, nb,
                                                                                             (int)m);
      LPB_KCHECK(lc);
      if (syrk_impl == 1)
        LPB_TRY(k_trailing_update_simple(lc, m, Mat, ldm, k0, nb));
      else
        LPB_TRY(k_trailing_update_dmma(lc, m, Mat, ldm, k0, nb));
    }
  }
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

int k_potrs(LaunchCtx& lc, int64_t m, const double* L, int64_t ldm, double* B, int nrhs) {
  LPB_TRY(configure_once());
  if (nrhs == 1) return potrs_impl<1>(lc, m, L, ldm, B);
  if (nrhs == 2) return potrs_impl<2>(lc, m, L, ldm, B);
  set_last_error("potrs: nrhs must be 1 or 2");
  return LPB_ERR_BAD_ARGUMENT;
}

}  // namespace lpb
