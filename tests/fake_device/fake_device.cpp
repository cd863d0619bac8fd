// TEST-ONLY host double of the device phase interface (never built into, linked to, or loaded by
// liblpb200.so).  It lets the CPU test-suite exercise lp_b200/csrc/ipm_driver.hpp -- the loop and
// scalar logic the product runs above the CUDA kernels -- against the oracle without a GPU.
// Plain scalar loops, unblocked lower Cholesky (the shape of the reference's default pure-Rust
// backend, newton_equations.rs:129-132,151-169).
#include <cmath>
#include <cstring>
#include <vector>

#include "../../lp_b200/csrc/ipm_driver.hpp"

namespace {

struct FakeDev {
  int64_t m, n;
  const double *A, *b, *c;
  std::vector<double> x, y, z, rP, rD, dinv, M, L, xs, r1, p, q, u, v, dx, dy, dz, rhs, tmp;
  bool have_pq = false;

  FakeDev(int64_t m_, int64_t n_, const double* A_, const double* b_, const double* c_)
      : m(m_), n(n_), A(A_), b(b_), c(c_), x(n_), y(m_), z(n_), rP(m_), rD(n_), dinv(n_), M(m_ * m_), L(m_ * m_),
        xs(n_), r1(n_), p(n_), q(m_), u(n_), v(m_), dx(n_), dy(m_), dz(n_), rhs(m_), tmp(m_) {}

  int blind_start() {
    for (auto& e : x) e = 1.0;
    for (auto& e : y) e = 0.0;
    for (auto& e : z) e = 1.0;
    return LPB_OK;
  }

  int residuals(double tau, double kappa, lpb_residual_scalars* o) {
    (void)kappa;
    double sp = 0, sd = 0, cx = 0, by = 0, xz = 0;
    for (int64_t i = 0; i < m; ++i) {
      double s = 0;
      for (int64_t j = 0; j < n; ++j) s += A[i * n + j] * x[j];
      rP[i] = b[i] * tau - s;
      sp += rP[i] * rP[i];
      by += b[i] * y[i];
    }
    for (int64_t j = 0; j < n; ++j) {
      double s = 0;
      for (int64_t i = 0; i < m; ++i) s += A[i * n + j] * y[i];
      rD[j] = c[j] * tau - s - z[j];
      sd += rD[j] * rD[j];
      cx += c[j] * x[j];
      xz += x[j] * z[j];
    }
    o->nrm_rp = std::sqrt(sp);
    o->nrm_rd = std::sqrt(sd);
    o->cx = cx;
    o->by = by;
    o->xz = xz;
    return LPB_OK;
  }

  int form_and_factor() {
    for (int64_t j = 0; j < n; ++j) dinv[j] = x[j] / z[j];
    for (int64_t i = 0; i < m; ++i)
      for (int64_t k = 0; k <= i; ++k) {
        double s = 0;
        for (int64_t j = 0; j < n; ++j) s += A[i * n + j] * dinv[j] * A[k * n + j];
        M[i * m + k] = s;
      }
    for (int64_t j = 0; j < m; ++j) {
      double s = M[j * m + j];
      for (int64_t l = 0; l < j; ++l) s -= L[j * m + l] * L[j * m + l];
      if (!(s > 0.0) || !std::isfinite(s)) return LPB_ERR_NUMERICAL_PROBLEM;
      const double d = std::sqrt(s);
      L[j * m + j] = d;
      for (int64_t i = j + 1; i < m; ++i) {
        double t = M[i * m + j];
        for (int64_t l = 0; l < j; ++l) t -= L[i * m + l] * L[j * m + l];
        L[i * m + j] = t / d;
      }
    }
    have_pq = false;
    return LPB_OK;
  }

  void chol_solve(std::vector<double>& r) {
    for (int64_t i = 0; i < m; ++i) {
      double s = r[i];
      for (int64_t l = 0; l < i; ++l) s -= L[i * m + l] * r[l];
      r[i] = s / L[i * m + i];
    }
    for (int64_t i = m - 1; i >= 0; --i) {
      double s = r[i];
      for (int64_t l = i + 1; l < m; ++l) s -= L[l * m + i] * r[l];
      r[i] = s / L[i * m + i];
    }
  }

  // newton_equations.rs:214-225
  void sym_solve(const double* r1v, const double* r2v, std::vector<double>& uo, std::vector<double>& vo) {
    for (int64_t i = 0; i < m; ++i) {
      double s = 0;
      for (int64_t j = 0; j < n; ++j) s += A[i * n + j] * (dinv[j] * r1v[j]);
      vo[i] = r2v[i] + s;
    }
    chol_solve(vo);
    for (int64_t j = 0; j < n; ++j) {
      double s = 0;
      for (int64_t i = 0; i < m; ++i) s += A[i * n + j] * vo[i];
      uo[j] = dinv[j] * (s - r1v[j]);
    }
  }

  int direction(const lpb_direction_in& in, double tau, double kappa, lpb_direction_out* o) {
    (void)tau;
    (void)kappa;
    const double gm = in.gamma * in.mu;
    if (!in.corrector) {
      for (int64_t j = 0; j < n; ++j) xs[j] = (x[j] * -1.0) * z[j] + gm;  // rhat.rs:32
    } else if (in.ip) {
      const double a2 = in.alpha * in.alpha;
      const double s = (1.0 - in.alpha) * in.gamma * in.mu;
      for (int64_t j = 0; j < n; ++j) xs[j] = (x[j] * -1.0) * z[j] - (dx[j] * dz[j]) * a2 + s;  // rhat.rs:54-55
    } else {
      for (int64_t j = 0; j < n; ++j) xs[j] = (x[j] * -1.0) * z[j] + gm - (dx[j] * dz[j]);  // rhat.rs:64
    }
    for (int64_t j = 0; j < n; ++j) r1[j] = rD[j] * in.eta - xs[j] / x[j];  // newton_equations.rs:188
    for (int64_t i = 0; i < m; ++i) tmp[i] = rP[i] * in.eta;
    if (!have_pq) {
      sym_solve(c, b, p, q);
      have_pq = true;
    }
    sym_solve(r1.data(), tmp.data(), u, v);
    o->cu = o->bv = o->cp = o->bq = 0;
    o->nan_pq = 0;
    for (int64_t j = 0; j < n; ++j) {
      o->cu += c[j] * u[j];
      o->cp += c[j] * p[j];
      if (std::isnan(p[j])) o->nan_pq = 1;
    }
    for (int64_t i = 0; i < m; ++i) {
      o->bv += b[i] * v[i];
      o->bq += b[i] * q[i];
      if (std::isnan(q[i])) o->nan_pq = 1;
    }
    return LPB_OK;
  }

  int assemble_delta(double d_tau, double axz[2]) {
    double ax = 1.0, az = 1.0;
    for (int64_t j = 0; j < n; ++j) {
      dx[j] = u[j] + p[j] * d_tau;             // delta.rs:33
      dz[j] = (xs[j] - z[j] * dx[j]) / x[j];   // delta.rs:37
      if (dx[j] < 0.0) ax = std::fmin(ax, x[j] / -dx[j]);
      if (dz[j] < 0.0) az = std::fmin(az, z[j] / -dz[j]);
    }
    for (int64_t i = 0; i < m; ++i) dy[i] = v[i] + q[i] * d_tau;  // delta.rs:34
    axz[0] = ax;
    axz[1] = az;
    return LPB_OK;
  }

  int do_step(double alpha, int ip) {
    for (int64_t j = 0; j < n; ++j) {
      x[j] = x[j] + dx[j] * alpha;
      z[j] = z[j] + dz[j] * alpha;
      if (ip) {
        x[j] = std::fmax(x[j], 1.0);
        z[j] = std::fmax(z[j], 1.0);
      }
    }
    for (int64_t i = 0; i < m; ++i) y[i] = y[i] + dy[i] * alpha;
    return LPB_OK;
  }
};

}  // namespace

extern "C" int fake_solve(int64_t m, int64_t n, const double* A, const double* b, const double* c, double c0,
                          const lpb_options* opts, double* x_out, double* fun, int64_t* iterations,
                          double* trace, int64_t max_rows, int64_t* n_rows) {
  int rc = lpb::options_validate(opts);
  if (rc != LPB_OK) return rc;
  FakeDev dev(m, n, A, b, c);
  lpb::SolveOutput out;
  rc = lpb::solve_normal_form(dev, *opts, n, c0, &out);
  if (iterations) *iterations = out.iterations;
  if (rc == LPB_OK || rc == LPB_ERR_ITERATION_LIMIT_EXCEEDED) {
    double f = 0;
    for (int64_t j = 0; j < n; ++j) {
      x_out[j] = dev.x[j] / out.tau;
      f += c[j] * x_out[j];
    }
    if (fun) *fun = f + c0;
  }
  int64_t rows = 0;
  for (; rows < (int64_t)out.trace.size() && rows < max_rows; ++rows)
    std::memcpy(trace + rows * LPB_TRACE_COLS, out.trace[rows].v, sizeof(double) * LPB_TRACE_COLS);
  if (n_rows) *n_rows = rows;
  return rc;
}

extern "C" void fake_options_default(lpb_options* o) { lpb::options_default(o); }
