// panel_factor.cuh -- the serial heart of every in-CTA Cholesky of this library: ONE warp factors a 16 x 16 diagonal
// sub-block held in shared memory and inverts it in the same 16 pivot steps.  Used by potf2_inv_kernel (cholesky.cu:
// the 128 x 128 panel kernel of K2) and by the batched interior-point kernel (batched.cu: K6, one LP per CTA).
// Replaces, together with its callers, LAPACK potrf behind /root/reference/src/solvers/interior_point/
// newton_equations.rs:129-132 (`cholesky_into`) / :87-90.
#pragma once

#include <cstdint>

namespace lpb {

// 1 / sqrt(a) for a pivot: the hardware seed (MUFU.RSQ64H: it reads the high word of the double, ~2^-22 relative over
// the whole FP64 exponent range, no conversions or range branch) and ONE third-order step  y (1 + e/2 + 3 e^2 / 8),
// e = 1 - a y^2  (truncation 5 e^3 / 16 ~ 2^-67; the result is within 1 ulp) -- four dependent FP64 operations behind
// the seed instead of the six of two Newton steps (the library's rsqrt: seed + Newton + special-case handling, ~20
// instructions); an FP64 operation costs ~20 cycles of latency here and this chain runs once per pivot.
// Non-positive / non-finite pivots come out as NaN / +inf / NaN, which is how factor_sub16 detects them; denormal
// pivots are flushed to zero by the seed, i.e. reported as bad.
__device__ __forceinline__ double pivot_rsqrt_refine(double a, double y) {
  const double e = fma(-(a * y), y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
}
__device__ __forceinline__ double pivot_rsqrt(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  return pivot_rsqrt_refine(a, y);
}

constexpr int kPfSB = 16;  // sub-block size
// scratch (doubles) behind `cbuf`: two pivot-column buffers of 16 + 16 padding, then a 16 x XDP dump for the factor lanes
__host__ __device__ constexpr int panel_factor_scratch(int xdp) { return 4 * kPfSB + kPfSB * xdp; }

// All 32 lanes of ONE warp.  Lanes 0..15 hold row i of the sub-block, lanes 16..31 the running sums of the forward
// substitution L x = e_i for column i of X = inv(L_dd); both need exactly column j of L at pivot step j (broadcast
// through `cbuf`) and then the same FMA.  `srow` = address of entry (i, 0) of the sub-block for lane i (both halves
// of the warp point at row i = lane & 15).  Results: L over the lower triangle of the sub-block, X into xd[r * XDP + c]
// (zero above its diagonal), 1 / L[i][i] into *rd (factor lanes, if rd != nullptr), and with WRITE_XT the strictly
// lower part of X transposed into the strictly upper triangle of the sub-block.  Returns a mask of the lanes (= rows)
// whose pivot was non-positive or non-finite (the first set bit is the first bad pivot; later ones are poisoned).
//
// The pivot loop is ROLLED: a panel kernel runs once per launch on an SM whose instruction cache is cold, and
// straight-line code that is executed once costs a cache-line fetch (~200 cycles) per eight instructions -- the fully
// unrolled version of potf2_inv_kernel spent 18.8 k cycles on its first sub-block and 4.2 k on the later ones
// (profiles/potf2_phases_before_r02.txt).  To keep the row in registers under a rolled loop it SHIFTS: at pivot j, v[k]
// is the entry of column j + k, the update  v[k] <- v[k+1] - l * L[j+1+k][j]  moves it down by one for free, and v[0]
// is always the pivot column's entry.  Entries past the sub-block's last column read the padding of the column buffer
// and are never consumed.  The body is branch-free (a divergent if / else would fence the rsqrt chain off from the
// column update): the inverse lanes' copy of l lands in the padding half of the column buffer, the factor lanes'
// copy of X in the dump.
template <int XDP, bool WRITE_XT>
__device__ __forceinline__ unsigned factor_sub16(double* srow, double* xd, double* cbuf, double* rd, int lane) {
  constexpr int SB = kPfSB;
  const unsigned full = 0xffffffffu;
  const int i = lane & 15;
  const bool inv_lane = lane >= SB;
  // One update rule for both halves of the warp.  Factor lanes: v = running a[i][.].  Inverse lanes: v = running
  // e_i[.] - sum_{l<.} L[.][l] x_l  (the right-hand side of L x = e_i after eliminating x_0 .. x_{j-1}).  At pivot j
  // both form  l = v[0] / sqrt(a_jj)  -- L[i][j] resp. x_j, and for lane j itself l = a_jj / sqrt(a_jj) = L[j][j] --
  // and both subtract  l * L[k][j]  from the entry of column k > j.
  double v[SB];
#pragma unroll
  for (int k = 0; k < SB; ++k) v[k] = inv_lane ? (k == i ? 1.0 : 0.0) : (k <= i ? srow[k] : 0.0);
  // the diagonal entry of this lane's row lives in its own register: the next pivot then depends only on
  // rsqrt -> l -> dg -= l * l, not on the shared-memory round trip of the column broadcast
  double dg = inv_lane ? 1.0 : srow[i];
  double* xcol = inv_lane ? xd + i : cbuf + 4 * SB + i;
  double myrd = 1.0;
  double rs = pivot_rsqrt(__shfl_sync(full, dg, 0));
#pragma unroll 1
  for (int j = 0; j < SB; ++j) {
    const double l = v[0] * rs;
    if (i == j) myrd = rs;
    dg = fma(-l, l, dg);  // only factor lanes i > j read it again (the inverse lanes' copy is never used)
    // the next pivot and its rsqrt start now: their latency overlaps the column update (j = 15 computes an unused rs)
    rs = pivot_rsqrt(__shfl_sync(full, dg, (j + 1) & (SB - 1)));
    double* cb = cbuf + (j & 1) * (2 * SB);
    cb[lane] = l;
    if (inv_lane ? (WRITE_XT && j > i) : (j <= i)) srow[j] = l;  // L[i][j] | X^T in the strictly upper triangle
    xcol[j * XDP] = (j >= i) ? l : 0.0;                          // X[r = j][c = i], zero above the diagonal
    __syncwarp();
    const double* cn = cb + j + 1;  // L[j+1 ..][j]; past row 15: padding, finite or not, never consumed
#pragma unroll
    for (int k = 0; k + 1 < SB; ++k) v[k] = fma(-l, cn[k], v[k + 1]);
  }
  // A non-positive / non-finite pivot leaves 1 / sqrt = NaN or +inf in its lane and NaN in every later one (the
  // arithmetic poisons the block by itself), so ONE test per sub-block finds the first bad pivot -- nothing on the
  // per-pivot chain.
  const bool bad = !inv_lane && !(myrd > 0.0 && myrd < __longlong_as_double(0x7ff0000000000000ll));
  if (!inv_lane && rd) *rd = myrd;
  return __ballot_sync(full, bad);
}

}  // namespace lpb
