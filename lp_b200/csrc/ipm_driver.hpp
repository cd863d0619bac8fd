// ipm_driver.hpp -- the host side of the interior-point iteration: loop + scalar logic.
//
// Device-agnostic: `Dev` supplies the phase calls (the same ones the C ABI exports in
// include/lpb200.h); this header owns everything the reference does with scalars.
// Restates, for the GPU phase split, /root/reference/src/solvers/interior_point/
//   mod.rs:199-240 (solve_normal_form), feasible_point.rs:53-72,110-165,
//   rhat.rs (scalar parts), delta.rs:29-38 (scalar parts), indicators.rs:37-83.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include "../../include/lpb200.h"

#if defined(__CUDACC__)
#define LPB_HD __host__ __device__
#else
#define LPB_HD
#endif

namespace lpb {

struct Indicators {  // indicators.rs:8-23
  double rho_p, rho_d, rho_A, rho_g, rho_mu, obj, bty;
};

enum class Status { Optimal, Infeasible, Unbounded, Unfinished };  // indicators.rs:85-90

// indicators.rs:66-83.  All comparisons are `<` / `>` so NaN indicators give Unfinished.
LPB_HD inline Status indicators_status(const Indicators& i, double tau, double kappa, double tol) {
  const bool tau_too_small = tau < tol * fmax(kappa, 1.0);
  const bool inf1 = (i.rho_p < tol && i.rho_d < tol && i.rho_g < tol) && tau_too_small;
  const bool inf2 = i.rho_mu < tol && tau_too_small;
  if (inf1 || inf2) return i.bty > tol ? Status::Infeasible : Status::Unbounded;
  if (i.rho_p < tol && i.rho_d < tol && i.rho_A < tol) return Status::Optimal;
  return Status::Unfinished;
}

struct InitialResiduals {  // residual.rs:5-10 evaluated at the blind start
  double rho_p, rho_d, rho_g, rho_mu;
};

// residual.rs:13-44 from the reduction scalars of one residual sweep.
LPB_HD inline void residual_values(const lpb_residual_scalars& rs, double tau, double kappa, int64_t n_total,
                            double* rho_p, double* rho_d, double* rho_g, double* rho_mu) {
  *rho_p = rs.nrm_rp;
  *rho_d = rs.nrm_rd;
  *rho_g = fabs(kappa + rs.cx - rs.by);
  *rho_mu = (rs.xz + tau * kappa) / static_cast<double>(n_total + 1);
}

// indicators.rs:37-55.
LPB_HD inline Indicators make_indicators(const lpb_residual_scalars& rs, const InitialResiduals& ini, double tau,
                                  double kappa, int64_t n_total, double c0) {
  Indicators out;
  double rp, rd, rg, rmu;
  residual_values(rs, tau, kappa, n_total, &rp, &rd, &rg, &rmu);
  out.obj = rs.cx / tau + c0;  // c.(x/tau) + c0 (display only)
  out.bty = rs.by;
  out.rho_A = fabs(rs.cx - rs.by) / (tau + fabs(rs.by));
  out.rho_p = rp / fmax(ini.rho_p, 1.0);
  out.rho_d = rd / fmax(ini.rho_d, 1.0);
  out.rho_g = rg / fmax(ini.rho_g, 1.0);
  out.rho_mu = rmu / ini.rho_mu;
  return out;
}

// Rust's f64::min: NaN-ignoring, like fmin.
LPB_HD inline double rmin(double a, double b) { return fmin(a, b); }

// feasible_point.rs:53-72; alpha_x / alpha_z come from the device ratio test (already min'ed with 1).
LPB_HD inline double step_size(double alpha_x, double alpha_z, double tau, double d_tau, double kappa, double d_kappa,
                        double alpha0) {
  const double alpha_tau = d_tau < 0.0 ? rmin(1.0, tau / -d_tau) : 1.0;
  const double alpha_kappa = d_kappa < 0.0 ? rmin(1.0, kappa / -d_kappa) : 1.0;
  return rmin(rmin(rmin(rmin(1.0, alpha_x), alpha_tau), alpha_z), alpha_kappa) * alpha0;
}

// feasible_point.rs:155-165.
LPB_HD inline double update_gamma(bool ip, double alpha) {
  if (ip) return 10.0;
  const double beta1 = 0.1;
  const double one_m = 1.0 - alpha;
  return (one_m * one_m) * rmin(beta1, one_m);
}

struct TraceRow {
  double v[LPB_TRACE_COLS];
};

// Plain-scalar result of the loop (usable from device code).
struct SolveScalars {
  int64_t iterations;
  double tau, kappa;
};

struct SolveOutput {
  int64_t iterations = 0;
  double tau = 1.0, kappa = 1.0;
  std::vector<TraceRow> trace;
};

// Host recorder: `disp` printing (mod.rs:208-211, :227-229; indicators.rs:25-33) + per-iteration trace.
struct HostRecorder {
  bool disp;
  std::vector<TraceRow>* trace;
  void start(const Indicators& i) {
    if (!disp) return;
    std::printf("alpha     \trho_p     \trho_d     \trho_g     \trho_mu    \tobj       \n");
    std::printf("1.00000000\t%.8f\t%.8f\t%.8f\t%.8f\t%8.3f\n", i.rho_p, i.rho_d, i.rho_g, i.rho_mu, i.obj);
  }
  void iteration(double alpha, const Indicators& i, double tau, double kappa, const double dbg[6]) {
    if (disp) std::printf("%.8f\t%.8f\t%.8f\t%.8f\t%.8f\t%8.3f\n", alpha, i.rho_p, i.rho_d, i.rho_g, i.rho_mu, i.obj);
    if (trace) {
      TraceRow row = {{alpha, i.rho_p, i.rho_d, i.rho_A, i.rho_g, i.rho_mu, i.obj, i.bty, tau, kappa, dbg[0], dbg[1],
                       dbg[2], dbg[3], dbg[4], dbg[5]}};
      trace->push_back(row);
    }
  }
};

// Device recorder (batched kernel): nothing to record.
struct NullRecorder {
  LPB_HD void start(const Indicators&) {}
  LPB_HD void iteration(double, const Indicators&, double, double, const double*) {}
};

// One Delta::compute worth of host scalars (delta.rs:29-32, :38).
LPB_HD inline void delta_scalars(double g_hat, double tk_hat, double tau, double kappa, const lpb_direction_out& d,
                          double* d_tau, double* d_kappa) {
  *d_tau = (g_hat + 1.0 / tau * tk_hat - (-d.cu + d.bv)) / (1.0 / tau * kappa + (-d.cp + d.bq));
  *d_kappa = 1.0 / tau * (tk_hat - kappa * *d_tau);
}

// solve_normal_form (mod.rs:199-240).  Returns an lpb status code; out->tau is the final tau so the
// caller can extract x / tau (mod.rs:231, :237-239).
//
// Dev must provide (all returning an lpb status code, 0 = ok):
//   blind_start(); residuals(tau, kappa, lpb_residual_scalars*); form_and_factor();
//   direction(const lpb_direction_in&, tau, kappa, lpb_direction_out*);
//   assemble_delta(d_tau, double alpha_xz[2]); do_step(alpha, ip);
template <class Dev, class Rec>
LPB_HD int solve_normal_form_rec(Dev& dev, const lpb_options& o, int64_t n_total, double c0, SolveScalars* out,
                                 Rec& rec) {
  int rc;
  double tau = 1.0, kappa = 1.0;  // feasible_point.rs:29-30
  out->iterations = 0;
  out->tau = tau;
  out->kappa = kappa;
  if ((rc = dev.blind_start()) != LPB_OK) return rc;

  lpb_residual_scalars rs;
  if ((rc = dev.residuals(tau, kappa, &rs)) != LPB_OK) return rc;
  InitialResiduals ini;  // feasible_point.rs:32
  residual_values(rs, tau, kappa, n_total, &ini.rho_p, &ini.rho_d, &ini.rho_g, &ini.rho_mu);

  Indicators ind = make_indicators(rs, ini, tau, kappa, n_total, c0);  // mod.rs:206
  rec.start(ind);

  bool ip = o.ip != 0;
  for (int64_t iteration = 1; iteration <= o.max_iter; ++iteration) {  // mod.rs:213
    // ---- get_delta (feasible_point.rs:110-152)
    double gamma = ip ? 1.0 : 0.0;                                   // :119
    double eta = ip ? 1.0 : 1.0 - gamma;                             // :120
    const double r_G = rs.cx - rs.by + kappa;                        // :124
    const double mu = (rs.xz + tau * kappa) / static_cast<double>(n_total + 1);  // :125

    if ((rc = dev.form_and_factor()) != LPB_OK) return rc;           // :127 (newton_equations.rs:48-64)

    // predictor (rhat.rs:17-35)
    lpb_direction_in din;
    lpb_direction_out dout;
    din.corrector = 0;
    din.ip = ip ? 1 : 0;
    din.eta = eta;
    din.gamma = gamma;
    din.mu = mu;
    din.alpha = 0.0;
    double g_hat = r_G * eta;                                        // rhat.rs:31
    double tk_hat = gamma * mu - tau * kappa;                        // rhat.rs:33
    if ((rc = dev.direction(din, tau, kappa, &dout)) != LPB_OK) return rc;
    if (dout.nan_pq) return LPB_ERR_NUMERICAL_PROBLEM;               // newton_equations.rs:190-194
    double d_tau, d_kappa;
    delta_scalars(g_hat, tk_hat, tau, kappa, dout, &d_tau, &d_kappa);
    double axz[2];
    if ((rc = dev.assemble_delta(d_tau, axz)) != LPB_OK) return rc;

    double alpha = step_size(axz[0], axz[1], tau, d_tau, kappa, d_kappa, 1.0);  // :134
    gamma = update_gamma(ip, alpha);                                 // :135
    eta = ip ? 1.0 : 1.0 - gamma;                                    // :136

    // corrector (rhat.rs:37-75)
    din.corrector = 1;
    din.eta = eta;
    din.gamma = gamma;
    din.alpha = alpha;
    g_hat = r_G * eta;
    if (ip) {  // rhat.rs:51-60
      const double alpha_2 = alpha * alpha;
      tk_hat = (1.0 - alpha) * gamma * mu - tau * kappa - alpha_2 * d_tau * d_kappa;
    } else {   // rhat.rs:62-66
      tk_hat = gamma * mu - tau * kappa - d_tau * d_kappa;
    }
    if ((rc = dev.direction(din, tau, kappa, &dout)) != LPB_OK) return rc;
    if (dout.nan_pq) return LPB_ERR_NUMERICAL_PROBLEM;
    delta_scalars(g_hat, tk_hat, tau, kappa, dout, &d_tau, &d_kappa);
    if ((rc = dev.assemble_delta(d_tau, axz)) != LPB_OK) return rc;

    // ---- step (mod.rs:216-223)
    alpha = ip ? 1.0 : step_size(axz[0], axz[1], tau, d_tau, kappa, d_kappa, o.alpha0);
    if ((rc = dev.do_step(alpha, ip ? 1 : 0)) != LPB_OK) return rc;
    tau = tau + d_tau * alpha;       // feasible_point.rs:80
    kappa = kappa + d_kappa * alpha; // :81
    if (ip) {                        // :92-93
      tau = fmax(tau, 1.0);
      kappa = fmax(kappa, 1.0);
    }
    ip = false;

    // ---- indicators (mod.rs:225-235)
    if ((rc = dev.residuals(tau, kappa, &rs)) != LPB_OK) return rc;
    ind = make_indicators(rs, ini, tau, kappa, n_total, c0);
    {  // parity-debugging scalars of the corrector: where two backends first part ways
      const double dbg[6] = {dout.cp, dout.bq, dout.cu, dout.bv, d_tau, d_kappa};
      rec.iteration(alpha, ind, tau, kappa, dbg);
    }
    out->iterations = iteration;
    out->tau = tau;
    out->kappa = kappa;
    switch (indicators_status(ind, tau, kappa, o.tol)) {
      case Status::Optimal: return LPB_OK;
      case Status::Infeasible: return LPB_ERR_INFEASIBLE;
      case Status::Unbounded: return LPB_ERR_UNBOUNDED;
      case Status::Unfinished: break;
    }
  }
  out->tau = tau;
  out->kappa = kappa;
  return LPB_ERR_ITERATION_LIMIT_EXCEEDED;  // mod.rs:237-239
}

// Host form: records the trace into `out` and honours `disp`.
template <class Dev>
int solve_normal_form(Dev& dev, const lpb_options& o, int64_t n_total, double c0, SolveOutput* out) {
  out->trace.clear();
  HostRecorder rec{o.disp != 0, &out->trace};
  SolveScalars sc;
  const int rc = solve_normal_form_rec(dev, o, n_total, c0, &sc, rec);
  out->iterations = sc.iterations;
  out->tau = sc.tau;
  out->kappa = sc.kappa;
  return rc;
}

// InteriorPointBuilder::new / build (mod.rs:51-60, :118-128)
LPB_HD inline void options_default(lpb_options* o) {
  o->tol = 1e-8;
  o->disp = 0;
  o->ip = 1;
  o->solver_type = LPB_SOLVER_CHOLESKY;
  o->reserved = 0;
  o->alpha0 = 0.99995;
  o->max_iter = 1000;
}

LPB_HD inline int options_validate(const lpb_options* o) {
  if (!o) return LPB_ERR_BAD_ARGUMENT;
  if (o->alpha0 <= 0.0 || o->alpha0 >= 1.0) return LPB_ERR_INVALID_PARAMETER;
  if (o->tol <= 0.0) return LPB_ERR_INVALID_PARAMETER;
  if (o->solver_type != LPB_SOLVER_CHOLESKY) return LPB_ERR_UNSUPPORTED;
  if (o->max_iter < 0) return LPB_ERR_BAD_ARGUMENT;
  return LPB_OK;
}

}  // namespace lpb
