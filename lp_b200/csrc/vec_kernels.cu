// vec_kernels.cu -- K4 (sweeps over A) and K5 (fused vector passes + reductions), sm_100a.
//
// All of these are HBM-bound: coalesced double2 loads, warp-shuffle + block reductions,
// deterministic two-stage grid reductions (no floating-point atomics), as few launches as the
// dependency chain of one IPM iteration allows (5 sweeps over A).
#include <cstdarg>

#include "kernels.hpp"

namespace lpb {

// ------------------------------------------------------------------ last-error plumbing
static thread_local char g_last_error[512] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_last_error; }

#define LPB_LAUNCH_CHECK(lc)                                                              \
  do {                                                                                    \
    (lc).launches++;                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) {                                                             \
      set_last_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return LPB_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

constexpr int kVecThreads = 256;

// one column per thread: the epilogues that fold the row-chunk partials of A^T v are latency-bound per column
static inline int vec_blocks_thin(int64_t n) {
  int64_t b = ceil_div(n, (int64_t)kVecThreads);
  if (b < 1) b = 1;
  if (b > kMaxRedBlocks) b = kMaxRedBlocks;
  return (int)b;
}
static inline int vec_blocks(int64_t n) {
  int64_t b = ceil_div(n, (int64_t)kVecThreads * 4);
  if (b < 1) b = 1;
  if (b > kMaxRedBlocks) b = kMaxRedBlocks;
  return (int)b;
}

// ------------------------------------------------------------------ device reduction helpers
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide reduce of up to NV values held per thread; thread 0 writes partials[val*kMaxRedBlocks+blk].
template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], const int (&op)[NV], double* partials,
                                                   int val_base) {
  __shared__ double sm[NV][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    double r = op[k] == kRedMin ? warp_min(v[k]) : warp_sum(v[k]);
    if (lane == 0) sm[k][warp] = r;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      double r = lane < nwarp ? sm[k][lane] : (op[k] == kRedMin ? INFINITY : 0.0);
      r = op[k] == kRedMin ? warp_min(r) : warp_sum(r);
      if (lane == 0) partials[(size_t)(val_base + k) * kMaxRedBlocks + blockIdx.x] = r;
    }
  }
}

struct RedSpecDev {
  int nvals;
  int nblocks[kMaxRedVals];
  int op[kMaxRedVals];
};

__global__ void finalize_reduce_kernel(const double* __restrict__ partials, RedSpecDev spec,
                                       double* __restrict__ out) {
  // one warp per value; fixed fold order -> deterministic
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= spec.nvals) return;
  const int nb = spec.nblocks[warp];
  const bool is_min = spec.op[warp] == kRedMin;
  double r = is_min ? INFINITY : 0.0;
  for (int i = lane; i < nb; i += 32) {
    const double p = partials[(size_t)warp * kMaxRedBlocks + i];
    r = is_min ? fmin(r, p) : r + p;
  }
  r = is_min ? warp_min(r) : warp_sum(r);
  if (lane == 0) out[warp] = r;
}

int reduce_finalize(LaunchCtx& lc, const RedSpec& spec) {
  RedSpecDev d;
  d.nvals = spec.nvals;
  for (int i = 0; i < kMaxRedVals; ++i) {
    d.nblocks[i] = spec.nblocks[i];
    d.op[i] = spec.op[i];
  }
  finalize_reduce_kernel<<<1, 32 * kMaxRedVals, 0, lc.stream>>>(lc.red_partials, d, lc.red_out);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

int fetch_scalars(LaunchCtx& lc, int nvals) {
  (void)nvals;  // always the whole line: the fault word rides along behind the kMaxRedVals results
  LPB_CUDA(cudaMemcpyAsync(lc.red_host, lc.red_out, sizeof(double) * (kMaxRedVals + 1), cudaMemcpyDeviceToHost,
                           lc.stream));
  LPB_CUDA(cudaStreamSynchronize(lc.stream));
  unsigned long long fault;
  std::memcpy(&fault, lc.red_host + kMaxRedVals, sizeof(fault));
  if (fault != 0ull) {
    set_last_error("a pipelined kernel gave up waiting for a hand-off from another CTA (fault word %llu)", fault);
    return LPB_ERR_CUDA;
  }
  return LPB_OK;
}

// ------------------------------------------------------------------ simple elementwise
__global__ void fill_kernel(double* __restrict__ p, int64_t n, double v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    p[i] = v;
}
int k_fill(LaunchCtx& lc, double* p, int64_t n, double v) {
  if (n <= 0) return LPB_OK;
  fill_kernel<<<vec_blocks(n), kVecThreads, 0, lc.stream>>>(p, n, v);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

int k_copy(LaunchCtx& lc, double* dst, const double* src, int64_t n) {
  if (n <= 0) return LPB_OK;
  LPB_CUDA(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToDevice, lc.stream));
  return LPB_OK;
}

__global__ void dinv_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ z,
                            double* __restrict__ dinv) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dinv[i] = x[i] / z[i];  // newton_equations.rs:54
}
int k_dinv(LaunchCtx& lc, int64_t n, const double* x, const double* z, double* dinv) {
  dinv_kernel<<<vec_blocks(n), kVecThreads, 0, lc.stream>>>(n, x, z, dinv);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

template <int MODE>
__global__ void rhat_kernel(int64_t n, double eta, double gm, double a2, double s, const double* __restrict__ x,
                            const double* __restrict__ z, const double* __restrict__ rD,
                            const double* __restrict__ dx, const double* __restrict__ dz, double* __restrict__ xs,
                            double* __restrict__ r1) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double xi = x[i];
    const double mxz = (xi * -1.0) * z[i];
    double v;
    if (MODE == 0) {
      v = mxz + gm;                          // rhat.rs:32
    } else if (MODE == 1) {
      v = mxz - (dx[i] * dz[i]) * a2 + s;    // rhat.rs:54-55
    } else {
      v = mxz + gm - (dx[i] * dz[i]);        // rhat.rs:64
    }
    xs[i] = v;
    r1[i] = rD[i] * eta - v / xi;            // newton_equations.rs:188 (rhat.d - rhat.xs / x)
  }
}
int k_rhat(LaunchCtx& lc, int64_t n, int mode, double eta, double gm, double a2, double s, const double* x,
           const double* z, const double* rD, const double* dx, const double* dz, double* xs, double* r1) {
  const int nb = vec_blocks(n);
  if (mode == 0)
    rhat_kernel<0><<<nb, kVecThreads, 0, lc.stream>>>(n, eta, gm, a2, s, x, z, rD, dx, dz, xs, r1);
  else if (mode == 1)
    rhat_kernel<1><<<nb, kVecThreads, 0, lc.stream>>>(n, eta, gm, a2, s, x, z, rD, dx, dz, xs, r1);
  else
    rhat_kernel<2><<<nb, kVecThreads, 0, lc.stream>>>(n, eta, gm, a2, s, x, z, rD, dx, dz, xs, r1);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

__global__ void assemble_delta_n_kernel(int64_t n, double d_tau, const double* __restrict__ u,
                                        const double* __restrict__ p, const double* __restrict__ xs,
                                        const double* __restrict__ x, const double* __restrict__ z,
                                        double* __restrict__ dx, double* __restrict__ dz,
                                        double* __restrict__ partials, int val_base) {
  double v[2] = {1.0, 1.0};  // fold(F::one(), min)  feasible_point.rs:61-62
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double xi = x[i], zi = z[i];
    const double dxi = u[i] + p[i] * d_tau;       // delta.rs:33
    const double dzi = (xs[i] - zi * dxi) / xi;   // delta.rs:37
    dx[i] = dxi;
    dz[i] = dzi;
    if (dxi < 0.0) v[0] = fmin(v[0], xi / -dxi);
    if (dzi < 0.0) v[1] = fmin(v[1], zi / -dzi);
  }
  const int op[2] = {kRedMin, kRedMin};
  block_reduce_store<2>(v, op, partials, val_base);
}
int k_assemble_delta_n(LaunchCtx& lc, int64_t n, double d_tau, const double* u, const double* p, const double* xs,
                       const double* x, const double* z, double* dx, double* dz, int val_base, int* nblocks) {
  const int nb = vec_blocks(n);
  assemble_delta_n_kernel<<<nb, kVecThreads, 0, lc.stream>>>(n, d_tau, u, p, xs, x, z, dx, dz, lc.red_partials,
                                                             val_base);
  LPB_LAUNCH_CHECK(lc);
  *nblocks = nb;
  return LPB_OK;
}

__global__ void assemble_delta_m_kernel(int64_t m, double d_tau, const double* __restrict__ v,
                                        const double* __restrict__ q, double* __restrict__ dy) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x)
    dy[i] = v[i] + q[i] * d_tau;  // delta.rs:34
}
int k_assemble_delta_m(LaunchCtx& lc, int64_t m, double d_tau, const double* v, const double* q, double* dy) {
  assemble_delta_m_kernel<<<vec_blocks(m), kVecThreads, 0, lc.stream>>>(m, d_tau, v, q, dy);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

__global__ void step_kernel(int64_t n, double alpha, int clamp, double* __restrict__ x,
                            const double* __restrict__ dx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double v = x[i] + dx[i] * alpha;  // feasible_point.rs:77-79
    if (clamp) v = fmax(v, 1.0);      // :89-91
    x[i] = v;
  }
}
int k_step(LaunchCtx& lc, int64_t n, double alpha, int clamp, double* x, const double* dx) {
  step_kernel<<<vec_blocks(n), kVecThreads, 0, lc.stream>>>(n, alpha, clamp, x, dx);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

__global__ void extract_x_kernel(int64_t n, double tau, const double* __restrict__ x, const double* __restrict__ c,
                                 double* __restrict__ xo, double* __restrict__ partials, int val) {
  double v[1] = {0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double t = x[i] / tau;  // mod.rs:231
    xo[i] = t;
    v[0] += c[i] * t;             // linear_program.rs:62
  }
  const int op[1] = {kRedSum};
  block_reduce_store<1>(v, op, partials, val);
}
int k_extract_x(LaunchCtx& lc, int64_t n, double tau, const double* x, const double* c, double* xo, int val,
                int* nblocks) {
  const int nb = vec_blocks(n);
  extract_x_kernel<<<nb, kVecThreads, 0, lc.stream>>>(n, tau, x, c, xo, lc.red_partials, val);
  LPB_LAUNCH_CHECK(lc);
  *nblocks = nb;
  return LPB_OK;
}

// ------------------------------------------------------------------ m-side epilogues
__global__ void resid_p_kernel(int64_t m, double tau, const double* __restrict__ b, const double* __restrict__ t,
                               const double* __restrict__ y, double* __restrict__ rP, double* __restrict__ partials,
                               int val_base) {
  double v[2] = {0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const double bi = b[i];
    const double r = bi * tau - t[i];  // feasible_point.rs:122 / residual.rs:23
    rP[i] = r;
    v[0] += r * r;
    v[1] += bi * y[i];
  }
  const int op[2] = {kRedSum, kRedSum};
  block_reduce_store<2>(v, op, partials, val_base);
}
int k_resid_p(LaunchCtx& lc, int64_t m, double tau, const double* b, const double* t, const double* y, double* rP,
              int val_base, int* nblocks) {
  const int nb = vec_blocks(m);
  resid_p_kernel<<<nb, kVecThreads, 0, lc.stream>>>(m, tau, b, t, y, rP, lc.red_partials, val_base);
  LPB_LAUNCH_CHECK(lc);
  *nblocks = nb;
  return LPB_OK;
}

__global__ void sym_fwd_rhs_kernel(int64_t m, double eta, const double* __restrict__ rP,
                                   const double* __restrict__ b, const double* __restrict__ t0,
                                   const double* __restrict__ t1, double* __restrict__ rhs0,
                                   double* __restrict__ rhs1, int with_pq) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    rhs0[i] = rP[i] * eta + t0[i];           // r2 + A (Dinv∘r1), r2 = rhat.p = r_P*eta
    if (with_pq) rhs1[i] = b[i] + t1[i];     // r2 = b
  }
}
// Iterative refinement of the sym_solve (see CudaDev::direction): residuals of the normal equations
// written with the already computed u = Dinv(A^T v - r1):  rho0 = rP*eta - A u,  rho1 = b - A p.
__global__ void refine_rhs_kernel(int64_t m, double eta, const double* __restrict__ rP, const double* __restrict__ b,
                                  const double* __restrict__ au, const double* __restrict__ ap,
                                  double* __restrict__ rho0, double* __restrict__ rho1, int with_pq) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    rho0[i] = rP[i] * eta - au[i];
    if (with_pq) rho1[i] = b[i] - ap[i];
  }
}
int k_refine_rhs(LaunchCtx& lc, int64_t m, double eta, const double* rP, const double* b, const double* au,
                 const double* ap, double* rho0, double* rho1, int with_pq) {
  refine_rhs_kernel<<<vec_blocks(m), kVecThreads, 0, lc.stream>>>(m, eta, rP, b, au, ap, rho0, rho1, with_pq);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}
// y += x
__global__ void axpy1_kernel(int64_t count, const double* __restrict__ x, double* __restrict__ y) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    y[i] += x[i];
}
int k_add_inplace(LaunchCtx& lc, int64_t count, const double* x, double* y) {
  if (count <= 0) return LPB_OK;
  axpy1_kernel<<<vec_blocks(count), kVecThreads, 0, lc.stream>>>(count, x, y);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

int k_sym_fwd_rhs(LaunchCtx& lc, int64_t m, double eta, const double* rP, const double* b, const double* t0,
                  const double* t1, double* rhs0, double* rhs1, int with_pq) {
  sym_fwd_rhs_kernel<<<vec_blocks(m), kVecThreads, 0, lc.stream>>>(m, eta, rP, b, t0, t1, rhs0, rhs1, with_pq);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

__global__ void dots_m_kernel(int64_t m, const double* __restrict__ b, const double* __restrict__ v,
                              const double* __restrict__ q, int with_pq, double* __restrict__ partials,
                              int val_base) {
  double r[3] = {0.0, 0.0, 0.0};
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const double bi = b[i];
    r[0] += bi * v[i];
    if (with_pq) {
      const double qi = q[i];
      r[1] += bi * qi;
      r[2] += (qi != qi) ? 1.0 : 0.0;
    }
  }
  const int op[3] = {kRedSum, kRedSum, kRedSum};
  block_reduce_store<3>(r, op, partials, val_base);
}
int k_dots_m(LaunchCtx& lc, int64_t m, const double* b, const double* v, const double* q, int with_pq, int val_base,
             int* nblocks) {
  const int nb = vec_blocks(m);
  dots_m_kernel<<<nb, kVecThreads, 0, lc.stream>>>(m, b, v, q, with_pq, lc.red_partials, val_base);
  LPB_LAUNCH_CHECK(lc);
  *nblocks = nb;
  return LPB_OK;
}

// ------------------------------------------------------------------ K4: t = A w   (row-major A, warp per row)
constexpr int kGemvNWarps = 8;

template <int NRHS, bool SCALE>
__global__ void __launch_bounds__(kGemvNWarps * 32)
gemv_n_kernel(int64_t m, int64_t n, const double* __restrict__ A, int64_t lda, const double* __restrict__ dinv,
              const double* __restrict__ w0, const double* __restrict__ w1, double* __restrict__ t0,
              double* __restrict__ t1) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * kGemvNWarps + warp;
  if (row >= m) return;
  const double2* __restrict__ Ar = reinterpret_cast<const double2*>(A + row * lda);
  const double2* __restrict__ W0 = reinterpret_cast<const double2*>(w0);
  const double2* __restrict__ W1 = reinterpret_cast<const double2*>(w1);
  const double2* __restrict__ D2 = reinterpret_cast<const double2*>(dinv);
  const int64_t n2 = n >> 1;
  double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 4
  for (int64_t j = lane; j < n2; j += 32) {
    const double2 a = __ldg(Ar + j);
    double2 w = W0[j];
    if (SCALE) {
      const double2 d = D2[j];
      w.x *= d.x;
      w.y *= d.y;
      if (NRHS == 2) {
        double2 ww = W1[j];
        ww.x *= d.x;
        ww.y *= d.y;
        acc1 += a.x * ww.x;
        acc1 += a.y * ww.y;
      }
    } else if (NRHS == 2) {
      const double2 ww = W1[j];
      acc1 += a.x * ww.x;
      acc1 += a.y * ww.y;
    }
    acc0 += a.x * w.x;
    acc0 += a.y * w.y;
  }
  if ((n & 1) && lane == 0) {
    const int64_t j = n - 1;
    const double a = A[row * lda + j];
    const double d = SCALE ? dinv[j] : 1.0;
    acc0 += a * (SCALE ? d * w0[j] : w0[j]);
    if (NRHS == 2) acc1 += a * (SCALE ? d * w1[j] : w1[j]);
  }
  acc0 = warp_sum(acc0);
  if (NRHS == 2) acc1 = warp_sum(acc1);
  if (lane == 0) {
    t0[row] = acc0;
    if (NRHS == 2) t1[row] = acc1;
  }
}

int k_gemv_n(LaunchCtx& lc, int64_t m, int64_t n, const double* A, int64_t lda, const double* dinv, const double* w0,
             const double* w1, double* t0, double* t1, int nrhs) {
  if ((lda & 1) || (reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(w0) & 15) ||
      (dinv && (reinterpret_cast<uintptr_t>(dinv) & 15)) || (nrhs == 2 && (reinterpret_cast<uintptr_t>(w1) & 15))) {
    set_last_error("gemv_n: A/w/dinv must be 16-byte aligned and lda even");
    return LPB_ERR_BAD_ARGUMENT;
  }
  const int nb = (int)ceil_div(m, kGemvNWarps);
  const dim3 block(kGemvNWarps * 32);
  if (nrhs == 2) {
    if (dinv)
      gemv_n_kernel<2, true><<<nb, block, 0, lc.stream>>>(m, n, A, lda, dinv, w0, w1, t0, t1);
    else
      gemv_n_kernel<2, false><<<nb, block, 0, lc.stream>>>(m, n, A, lda, dinv, w0, w1, t0, t1);
  } else {
    if (dinv)
      gemv_n_kernel<1, true><<<nb, block, 0, lc.stream>>>(m, n, A, lda, dinv, w0, w0, t0, t0);
    else
      gemv_n_kernel<1, false><<<nb, block, 0, lc.stream>>>(m, n, A, lda, dinv, w0, w0, t0, t0);
  }
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

// ------------------------------------------------------------------ K4: s = A^T v  (row chunks -> partials)
constexpr int kGemvTThreads = 128;
constexpr int kGemvTMaxRows = 1024;  // rows per chunk staged in smem (8 KB per rhs)
constexpr int kGemvTMaxChunks = 64;

template <int NRHS>
__global__ void __launch_bounds__(kGemvTThreads)
gemv_t_kernel(int64_t m, int64_t n, const double* __restrict__ A, int64_t lda, const double* __restrict__ v0,
              const double* __restrict__ v1, double* __restrict__ partials, int64_t n_pad, int rows_per_chunk) {
  __shared__ double sv[NRHS][kGemvTMaxRows];
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  int64_t r1 = r0 + rows_per_chunk;
  if (r1 > m) r1 = m;
  const int nr = (int)(r1 - r0);
  for (int i = threadIdx.x; i < nr; i += blockDim.x) {
    sv[0][i] = v0[r0 + i];
    if (NRHS == 2) sv[1][i] = v1[r0 + i];
  }
  __syncthreads();
  const int64_t j2 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // double2 column index
  const int64_t j = j2 * 2;
  if (j >= n) return;
  double ax0 = 0.0, ay0 = 0.0, ax1 = 0.0, ay1 = 0.0;
  if (j + 1 < n) {
    const double2* __restrict__ Ap = reinterpret_cast<const double2*>(A + r0 * lda) + j2;
    const int64_t ld2 = lda >> 1;
#pragma unroll 8
    for (int i = 0; i < nr; ++i) {
      const double2 a = __ldg(Ap + (int64_t)i * ld2);
      const double s0 = sv[0][i];
      ax0 += a.x * s0;
      ay0 += a.y * s0;
      if (NRHS == 2) {
        const double s1 = sv[1][i];
        ax1 += a.x * s1;
        ay1 += a.y * s1;
      }
    }
  } else {  // last odd column
    const double* __restrict__ Ap = A + r0 * lda + j;
    for (int i = 0; i < nr; ++i) {
      const double a = __ldg(Ap + (int64_t)i * lda);
      ax0 += a * sv[0][i];
      if (NRHS == 2) ax1 += a * sv[1][i];
    }
  }
  double* P0 = partials + ((size_t)blockIdx.y * NRHS + 0) * n_pad + j;
  P0[0] = ax0;
  P0[1] = ay0;  // n_pad is even, so j+1 < n_pad always
  if (NRHS == 2) {
    double* P1 = partials + ((size_t)blockIdx.y * NRHS + 1) * n_pad + j;
    P1[0] = ax1;
    P1[1] = ay1;
  }
}

static inline void gemv_t_shape(int64_t m, int64_t n, int* nchunks, int* rows_per_chunk, int* colblocks) {
  const int64_t cb = ceil_div(ceil_div(n, 2), kGemvTThreads);
  int64_t want = ceil_div((int64_t)kNumSMs * 8, cb);  // ~8 CTAs of 128 threads per SM
  if (want < 1) want = 1;
  if (want > kGemvTMaxChunks) want = kGemvTMaxChunks;
  int64_t rpc = ceil_div(m, want);
  if (rpc < 8) rpc = 8;
  if (rpc > kGemvTMaxRows) rpc = kGemvTMaxRows;
  int64_t nc = ceil_div(m, rpc);
  if (nc > kGemvTMaxChunks) {  // m very large: more chunks than the cap -> grow rows per chunk is impossible
    nc = ceil_div(m, kGemvTMaxRows);
    rpc = kGemvTMaxRows;
  }
  *nchunks = (int)nc;
  *rows_per_chunk = (int)rpc;
  *colblocks = (int)cb;
}

// Workspace bound for ANY column count n' <= n (the sweeps run over the dense leading columns only, and fewer
// columns mean more row chunks): chunks never exceed max(kGemvTMaxChunks, ceil(m / kGemvTMaxRows)).
int64_t gemv_t_partials_doubles(int64_t m, int64_t n) {
  int64_t nc = ceil_div(m, kGemvTMaxRows);
  if (nc < kGemvTMaxChunks) nc = kGemvTMaxChunks;
  return nc * 2 * round_up(n, 2);
}

int k_gemv_t_partials(LaunchCtx& lc, int64_t m, int64_t n, const double* A, int64_t lda, const double* v0,
                      const double* v1, int nrhs, int* nchunks) {
  if ((lda & 1) || (reinterpret_cast<uintptr_t>(A) & 15)) {
    set_last_error("gemv_t: A must be 16-byte aligned and lda even");
    return LPB_ERR_BAD_ARGUMENT;
  }
  int nc, rpc, cb;
  gemv_t_shape(m, n, &nc, &rpc, &cb);
  const int64_t n_pad = round_up(n, 2);
  if ((int64_t)nc * nrhs * n_pad > lc.gemv_partials_cap) {
    set_last_error("gemv_t: partials workspace too small (%lld > %lld)", (long long)((int64_t)nc * nrhs * n_pad),
                   (long long)lc.gemv_partials_cap);
    return LPB_ERR_BAD_ARGUMENT;
  }
  const dim3 grid(cb, nc);
  if (nrhs == 2)
    gemv_t_kernel<2><<<grid, kGemvTThreads, 0, lc.stream>>>(m, n, A, lda, v0, v1, lc.gemv_partials, n_pad, rpc);
  else
    gemv_t_kernel<1><<<grid, kGemvTThreads, 0, lc.stream>>>(m, n, A, lda, v0, v0, lc.gemv_partials, n_pad, rpc);
  LPB_LAUNCH_CHECK(lc);
  *nchunks = nc;
  return LPB_OK;
}

__device__ __forceinline__ double sum_chunks(const double* __restrict__ partials, int nchunks, int nrhs, int k,
                                             int64_t n_pad, int64_t j) {
  // four interleaved partial sums in a fixed order: one chain of nchunks dependent L2 loads + adds per column was what
  // these epilogues cost at C2 (50 chunks: 62 us for sym_back_kernel on 8 CTAs; launches_C2_r02_final.txt)
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  const double* pk = partials + (size_t)k * n_pad + j;
  const size_t step = (size_t)nrhs * n_pad;
  int c = 0;
  for (; c + 3 < nchunks; c += 4) {
#pragma unroll
    for (int e = 0; e < 4; ++e) s[e] += pk[(size_t)(c + e) * step];
  }
  for (; c < nchunks; ++c) s[0] += pk[(size_t)c * step];
  return (s[0] + s[1]) + (s[2] + s[3]);
}

// (A^T v_k)[j] for the epilogues: columns below tc.nd come from the row-chunk partials (pitch round_up(nd, 2));
// the trailing singleton / zero columns (lpb_ctx::n_dense) are s_j * v_k[row_j] and never touch A.
__device__ __forceinline__ double at_dot(const double* __restrict__ partials, int nchunks, int nrhs, int k,
                                         const TailCols& tc, int64_t j) {
  if (j < tc.nd) return sum_chunks(partials, nchunks, nrhs, k, (tc.nd + 1) & ~(int64_t)1, j);
  const int r = tc.col_row[j];
  return r >= 0 ? tc.col_val[j] * (k ? tc.v1 : tc.v0)[r] : 0.0;
}

__global__ void gemv_t_raw_kernel(int64_t n, int64_t n_pad, int nchunks, int nrhs,
                                  const double* __restrict__ partials, double* __restrict__ out0,
                                  double* __restrict__ out1) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    out0[j] = sum_chunks(partials, nchunks, nrhs, 0, n_pad, j);
    if (nrhs == 2) out1[j] = sum_chunks(partials, nchunks, nrhs, 1, n_pad, j);
  }
}
int k_gemv_t_raw(LaunchCtx& lc, int64_t n, int nchunks, int nrhs, double* out0, double* out1) {
  gemv_t_raw_kernel<<<vec_blocks(n), kVecThreads, 0, lc.stream>>>(n, round_up(n, 2), nchunks, nrhs,
                                                                  lc.gemv_partials, out0, out1);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

__global__ void resid_d_kernel(int64_t n, TailCols tc, int nchunks, double tau,
                               const double* __restrict__ partials, const double* __restrict__ c,
                               const double* __restrict__ z, const double* __restrict__ x, double* __restrict__ rD,
                               double* __restrict__ red, int val_base) {
  double v[3] = {0.0, 0.0, 0.0};
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const double s = at_dot(partials, nchunks, 1, 0, tc, j);
    const double cj = c[j], zj = z[j], xj = x[j];
    const double r = cj * tau - s - zj;  // feasible_point.rs:123 / residual.rs:24-26
    rD[j] = r;
    v[0] += r * r;
    v[1] += cj * xj;
    v[2] += xj * zj;
  }
  const int op[3] = {kRedSum, kRedSum, kRedSum};
  block_reduce_store<3>(v, op, red, val_base);
}
int k_resid_d(LaunchCtx& lc, int64_t n, int nchunks, double tau, const double* c, const double* z, const double* x,
              double* rD, int val_base, int* nblocks, const TailCols* tail) {
  const int nb = vec_blocks_thin(n);
  const TailCols tc = tail ? *tail : TailCols{n, nullptr, nullptr, nullptr, nullptr};
  resid_d_kernel<<<nb, kVecThreads, 0, lc.stream>>>(n, tc, nchunks, tau, lc.gemv_partials, c, z, x, rD,
                                                    lc.red_partials, val_base);
  LPB_LAUNCH_CHECK(lc);
  *nblocks = nb;
  return LPB_OK;
}

__global__ void sym_back_kernel(int64_t n, TailCols tc, int nchunks, int with_pq,
                                const double* __restrict__ partials, const double* __restrict__ dinv,
                                const double* __restrict__ r1, const double* __restrict__ c, double* __restrict__ u,
                                double* __restrict__ p, double* __restrict__ red, int val_base) {
  double v[3] = {0.0, 0.0, 0.0};
  const int nrhs = with_pq ? 2 : 1;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const double dj = dinv[j], cj = c[j];
    const double s0 = at_dot(partials, nchunks, nrhs, 0, tc, j);
    const double uj = dj * (s0 - r1[j]);  // newton_equations.rs:223
    u[j] = uj;
    v[0] += cj * uj;
    if (with_pq) {
      const double s1 = at_dot(partials, nchunks, nrhs, 1, tc, j);
      const double pj = dj * (s1 - cj);
      p[j] = pj;
      v[1] += cj * pj;
      v[2] += (pj != pj) ? 1.0 : 0.0;
    }
  }
  const int op[3] = {kRedSum, kRedSum, kRedSum};
  block_reduce_store<3>(v, op, red, val_base);
}
int k_sym_back(LaunchCtx& lc, int64_t n, int nchunks, int with_pq, const double* dinv, const double* r1,
               const double* c, double* u, double* p, int val_base, int* nblocks, const TailCols* tail) {
  const int nb = vec_blocks_thin(n);
  const TailCols tc = tail ? *tail : TailCols{n, nullptr, nullptr, nullptr, nullptr};
  sym_back_kernel<<<nb, kVecThreads, 0, lc.stream>>>(n, tc, nchunks, with_pq, lc.gemv_partials, dinv, r1,
                                                     c, u, p, lc.red_partials, val_base);
  LPB_LAUNCH_CHECK(lc);
  *nblocks = nb;
  return LPB_OK;
}

// ------------------------------------------------------------------ synthetic shard fill (multi-GPU config C5)
// Counter-based generator: element (row, col) of the m x n0 Gaussian block depends only on
// (seed, row, col), so any column sharding reproduces the same matrix.
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ double counter_normal(uint64_t seed, uint64_t row, uint64_t col) {
  const uint64_t k = splitmix64(seed ^ splitmix64(row * 0x100000001B3ull + 0x1234567ull) ^
                                splitmix64(col + 0xABCDEF0123ull));
  const uint64_t a = splitmix64(k), b = splitmix64(k ^ 0xD1B54A32D192ED03ull);
  const double u1 = ((a >> 11) + 1.0) * (1.0 / 9007199254740993.0);  // (0,1]
  const double u2 = (b >> 11) * (1.0 / 9007199254740992.0);          // [0,1)
  return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}
__device__ __forceinline__ double counter_uniform(uint64_t seed, uint64_t idx, double lo, double hi) {
  const uint64_t a = splitmix64(seed ^ splitmix64(idx + 0x5851F42D4C957F2Dull));
  return lo + (hi - lo) * ((a >> 11) * (1.0 / 9007199254740992.0));
}

// kind 0: U(lo, hi);  kind 1: N(0,1);  kind 2: -|N(0,1)| for idx < neg_below, N(0,1) above (the y0 of
// the SURVEY 8(d) generator: dual-feasible multipliers, non-positive on the inequality rows).
__global__ void fill_vec_kernel(double* __restrict__ v, int64_t count, int64_t idx0, uint64_t seed, int kind, double lo,
                                double hi, int64_t neg_below) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t gi = (uint64_t)(idx0 + i);
    double r;
    if (kind == 0) {
      r = counter_uniform(seed, gi, lo, hi);
    } else {
      r = counter_normal(seed, gi, 0x7777ull);
      if (kind == 2 && (int64_t)gi < neg_below) r = -fabs(r);
    }
    v[i] = r;
  }
}
int k_fill_vec(LaunchCtx& lc, double* v, int64_t count, int64_t idx0, uint64_t seed, int kind, double lo, double hi,
               int64_t neg_below) {
  if (count <= 0) return LPB_OK;
  fill_vec_kernel<<<vec_blocks(count), kVecThreads, 0, lc.stream>>>(v, count, idx0, seed, kind, lo, hi, neg_below);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

// slack columns of the slack form: A[r][c] = 1 if (global column) == n0 + r else 0, for global columns >= n0
__global__ void slack_identity_kernel(double* __restrict__ A, int64_t rows, int64_t cols, int64_t lda, int64_t col0,
                                      int64_t n0) {
  const int64_t total = rows * cols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / cols, c = idx - r * cols;
    const int64_t gc = col0 + c;
    if (gc >= n0) A[r * lda + c] = (gc - n0 == r) ? 1.0 : 0.0;
  }
}
int k_slack_identity(LaunchCtx& lc, double* A, int64_t rows, int64_t cols, int64_t lda, int64_t col0, int64_t n0) {
  if (rows <= 0 || cols <= 0) return LPB_OK;
  slack_identity_kernel<<<kNumSMs * 8, 256, 0, lc.stream>>>(A, rows, cols, lda, col0, n0);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

// out[i] = a[i] + (b ? b[i] : 0) for i < count_b, a[i] beyond (count_b <= count)
__global__ void add_vec_kernel(double* __restrict__ out, const double* __restrict__ a, const double* __restrict__ b,
                               int64_t count, int64_t count_b) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = a[i] + (i < count_b ? b[i] : 0.0);
}
int k_add_vec(LaunchCtx& lc, double* out, const double* a, const double* b, int64_t count, int64_t count_b) {
  if (count <= 0) return LPB_OK;
  add_vec_kernel<<<vec_blocks(count), kVecThreads, 0, lc.stream>>>(out, a, b, count, count_b);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

__global__ void fill_normal_kernel(double* __restrict__ A, int64_t rows, int64_t cols, int64_t lda, int64_t row0,
                                   int64_t col0, uint64_t seed) {
  const int64_t total = rows * cols;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / cols, c = idx - r * cols;
    A[r * lda + c] = counter_normal(seed, (uint64_t)(row0 + r), (uint64_t)(col0 + c));
  }
}
int k_fill_normal(LaunchCtx& lc, double* A, int64_t rows, int64_t cols, int64_t lda, int64_t row0, int64_t col0,
                  uint64_t seed) {
  if (rows <= 0 || cols <= 0) return LPB_OK;
  fill_normal_kernel<<<kNumSMs * 8, 256, 0, lc.stream>>>(A, rows, cols, lda, row0, col0, seed);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}


// Order-independent 64-bit checksum of a (rows x cols, leading dimension ld) block of doubles: sum over
// entries of bits * (odd multiplier of the position), wrapping.  Integer atomics, so the value does not
// depend on scheduling; replicas of the sharded path compare it to prove they hold identical bits.
__global__ void checksum_kernel(const double* __restrict__ v, int64_t rows, int64_t cols, int64_t ld, int lower_only,
                                unsigned long long* __restrict__ out) {
  unsigned long long h = 0;
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    if (lower_only && c > r) continue;
    const unsigned long long bits = (unsigned long long)__double_as_longlong(v[r * ld + c]);
    h += bits * (2ull * (unsigned long long)i + 0x9E3779B97F4A7C15ull);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
  if ((threadIdx.x & 31) == 0 && h) atomicAdd(out, h);
}
int k_checksum(LaunchCtx& lc, const double* v, int64_t rows, int64_t cols, int64_t ld, int lower_only,
               unsigned long long* out_dev) {
  LPB_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(unsigned long long), lc.stream));
  if (rows <= 0 || cols <= 0) return LPB_OK;
  checksum_kernel<<<kNumSMs * 8, 256, 0, lc.stream>>>(v, rows, cols, ld, lower_only, out_dev);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}


// ------------------------------------------------------------------ column structure of A (slack columns)
// The slack form [[A_ub, I], [A_eq, 0]] (linear_program.rs:145-156) ends in n_slack singleton columns.
// A singleton column s e_r contributes s^2 d_j to M[r][r] and nothing else, so the SYRK only has to
// contract over the dense leading columns.  One pass over A: per column the number of non-zeros, the
// (last) row holding one and its value.  Row chunks run in parallel; the counts meet in atomics, and
// `val` is only meaningful when the final count is exactly 1 (then exactly one thread stored it).
__global__ void col_structure_kernel(const double* __restrict__ A, int64_t m, int64_t n, int64_t lda,
                                     int rows_per_chunk, int* __restrict__ nnz, int* __restrict__ row,
                                     double* __restrict__ val) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
  int64_t r1 = r0 + rows_per_chunk;
  if (r1 > m) r1 = m;
  int cnt = 0, last = -1;
  double v = 0.0;
  const double* __restrict__ Ap = A + j;
#pragma unroll 4
  for (int64_t r = r0; r < r1; ++r) {
    const double a = __ldg(Ap + r * lda);
    if (a != 0.0) {  // NaN counts as a non-zero
      ++cnt;
      last = (int)r;
      v = a;
    }
  }
  if (cnt > 0) {
    atomicAdd(nnz + j, cnt);
    atomicMax(row + j, last);
    if (cnt == 1) val[j] = v;
  }
}
int k_col_structure(LaunchCtx& lc, const double* A, int64_t m, int64_t n, int64_t lda, int* nnz, int* row,
                    double* val) {
  if (m <= 0 || n <= 0) return LPB_OK;
  LPB_CUDA(cudaMemsetAsync(nnz, 0, sizeof(int) * (size_t)n, lc.stream));
  LPB_CUDA(cudaMemsetAsync(row, 0xff, sizeof(int) * (size_t)n, lc.stream));  // -1
  LPB_CUDA(cudaMemsetAsync(val, 0, sizeof(double) * (size_t)n, lc.stream));
  const int cb = (int)ceil_div(n, 128);
  int64_t chunks = ceil_div((int64_t)kNumSMs * 16, cb);
  if (chunks > 1024) chunks = 1024;
  if (chunks > m) chunks = m;
  const int rpc = (int)ceil_div(m, chunks);
  col_structure_kernel<<<dim3(cb, (unsigned)ceil_div(m, rpc)), 128, 0, lc.stream>>>(A, m, n, lda, rpc, nnz, row, val);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

// M[r][r] += s_r^2 * dinv[col[r]] for the rows that own a singleton column (col[r] >= 0, value s_r).
__global__ void diag_add_kernel(int64_t m, double* __restrict__ M, int64_t ldm, const int* __restrict__ col,
                                const double* __restrict__ val, const double* __restrict__ dinv) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  const int j = col[r];
  if (j >= 0) M[r * ldm + r] += (val[r] * val[r]) * dinv[j];
}
int k_diag_add(LaunchCtx& lc, int64_t m, double* M, int64_t ldm, const int* col, const double* val,
               const double* dinv) {
  if (m <= 0) return LPB_OK;
  diag_add_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, lc.stream>>>(m, M, ldm, col, val, dinv);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

__global__ void diag_shift_kernel(int64_t m, double* __restrict__ M, int64_t ldm, double delta) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  const double d = M[r * ldm + r];
  M[r * ldm + r] = d > 0.0 ? d * (1.0 + delta) : d + delta;
}
int k_diag_shift(LaunchCtx& lc, int64_t m, double* M, int64_t ldm, double delta) {
  if (m <= 0) return LPB_OK;
  diag_shift_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, lc.stream>>>(m, M, ldm, delta);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

// The singleton columns' share of t_k = A (dinv? dinv * w_k : w_k):  t_k[r] += s_r * (dinv[j] *) w_k[j], j = col[r].
__global__ void slack_add_kernel(int64_t m, const int* __restrict__ col, const double* __restrict__ val,
                                 const double* __restrict__ dinv, const double* __restrict__ w0,
                                 const double* __restrict__ w1, double* __restrict__ t0, double* __restrict__ t1,
                                 int nrhs) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  const int j = col[r];
  if (j < 0) return;
  const double sd = dinv ? val[r] * dinv[j] : val[r];
  t0[r] += sd * w0[j];
  if (nrhs == 2) t1[r] += sd * w1[j];
}
int k_slack_add(LaunchCtx& lc, int64_t m, const int* col, const double* val, const double* dinv, const double* w0,
                const double* w1, double* t0, double* t1, int nrhs) {
  if (m <= 0) return LPB_OK;
  slack_add_kernel<<<(unsigned)ceil_div(m, 256), 256, 0, lc.stream>>>(m, col, val, dinv, w0, w1, t0, t1, nrhs);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

// ------------------------------------------------------------------ lower-triangle pack / unpack (sharded contexts)
// Only the lower triangle of M is ever read, so the all-reduce that assembles M from the ranks' partial
// SYRKs moves the lower triangle alone: block row i (128 rows) contributes its columns [0, 128 (i + 1)),
// stored contiguously at offset 128 * 128 * i (i + 1) / 2 with pitch 128 (i + 1).  Half the NVLink bytes for two
// extra HBM passes over the triangle.
constexpr int kTriBlk = 128;
template <bool PACK>
__global__ void __launch_bounds__(256)
tri_pack_kernel(double* __restrict__ M, int64_t ldm, int64_t m, double* __restrict__ buf) {
  const int64_t i = blockIdx.y;                       // block row
  const int64_t r0 = i * kTriBlk;
  const int64_t rows = (m - r0) < kTriBlk ? (m - r0) : kTriBlk;
  int64_t width = (i + 1) * kTriBlk;
  if (width > m) width = m;
  const int64_t pitch = (i + 1) * kTriBlk;
  double* dst = buf + (int64_t)kTriBlk * kTriBlk * (i * (i + 1) / 2);
  const int64_t total = rows * width;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = e / width, cc = e - r * width;
    if (PACK)
      dst[r * pitch + cc] = M[(r0 + r) * ldm + cc];
    else
      M[(r0 + r) * ldm + cc] = dst[r * pitch + cc];
  }
}
int64_t tri_packed_doubles(int64_t m) {
  const int64_t T = ceil_div(m, kTriBlk);
  return (int64_t)kTriBlk * kTriBlk * (T * (T + 1) / 2);
}
int k_tri_pack(LaunchCtx& lc, double* M, int64_t ldm, int64_t m, double* buf, bool pack) {
  if (m <= 0) return LPB_OK;
  const dim3 grid(64, (unsigned)ceil_div(m, kTriBlk));
  if (pack) {
    // rows beyond m / columns beyond min(width, m) of the last blocks are never written: keep them defined
    tri_pack_kernel<true><<<grid, 256, 0, lc.stream>>>(M, ldm, m, buf);
  } else {
    tri_pack_kernel<false><<<grid, 256, 0, lc.stream>>>(M, ldm, m, buf);
  }
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

// Debug: compare two matrices bit for bit (lower triangle if lower_only); out = {count, min col, min row}.
__global__ void diff_kernel(const double* __restrict__ a, const double* __restrict__ b, int64_t rows, int64_t cols,
                            int64_t ld, int lower_only, unsigned long long* __restrict__ out) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols, c = i - r * cols;
    if (lower_only && c > r) continue;
    if (__double_as_longlong(a[r * ld + c]) != __double_as_longlong(b[r * ld + c])) {
      atomicAdd(out, 1ull);
      atomicMin(out + 1, (unsigned long long)c);
      atomicMin(out + 2, (unsigned long long)r);
    }
  }
}
int k_diff(LaunchCtx& lc, const double* a, const double* b, int64_t rows, int64_t cols, int64_t ld, int lower_only,
           unsigned long long* out_dev3) {
  const unsigned long long init[3] = {0ull, ~0ull, ~0ull};
  LPB_CUDA(cudaMemcpyAsync(out_dev3, init, sizeof(init), cudaMemcpyHostToDevice, lc.stream));
  LPB_CUDA(cudaStreamSynchronize(lc.stream));
  diff_kernel<<<kNumSMs * 8, 256, 0, lc.stream>>>(a, b, rows, cols, ld, lower_only, out_dev3);
  LPB_LAUNCH_CHECK(lc);
  return LPB_OK;
}

}  // namespace lpb
