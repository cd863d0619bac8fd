// ripped.hpp -- C++ host-side mirror of the reference's public API, above the C ABI.
//
// The reference is compiled (Rust) code and this image has no Rust toolchain, so the host side
// that the north star asks for ("the host drives the iteration loop and calls the kernels through
// a thin extern "C" FFI") is written in C++ here; rust/ holds the equivalent (unbuilt) Rust shim.
// Names and semantics follow /root/reference/src:
//   ripped::Problem / ProblemBuilder          linear_program.rs:24-169
//   ripped::InteriorPoint / InteriorPointBuilder   solvers/interior_point/mod.rs:41-197
//   ripped::EquationSolverType                solvers/interior_point/newton_equations.rs:36-46
//   ripped::OptimizeResult, ripped::Solver    solvers/mod.rs:12-49
//   ripped::LinearProgramError                error.rs:7-29
// Rust's Result<T, E> is `ripped::Result<T>`: `ok()`, `value()`, `error()`, `unwrap()` (throws).
//
// InteriorPoint::solve runs lp_b200/csrc/ipm_driver.hpp (the same loop lpb_solve runs inside the
// library) on a `CabiDevice`, i.e. it drives lpb_blind_start / lpb_residuals / lpb_form_and_factor /
// lpb_direction / lpb_assemble_delta / lpb_do_step itself.  Header-only; link with -llpb200.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/lpb200.h"
#include "../csrc/ipm_driver.hpp"

namespace ripped {

// ------------------------------------------------------------------ error.rs:7-29
struct LinearProgramError {
  enum Kind {
    Unconstrained = LPB_ERR_UNCONSTRAINED,
    NumericalProblem = LPB_ERR_NUMERICAL_PROBLEM,
    InvalidParameter = LPB_ERR_INVALID_PARAMETER,
    IncompatibleInputDimensions = LPB_ERR_INCOMPATIBLE_INPUT_DIMENSIONS,
    Infeasible = LPB_ERR_INFEASIBLE,
    Unbounded = LPB_ERR_UNBOUNDED,
    IterationLimitExceeded = LPB_ERR_ITERATION_LIMIT_EXCEEDED,
    Device = 100  // CUDA / NCCL / argument failure: no reference counterpart
  } kind;
  std::string message;            // InvalidParameter(&'static str) payload, or lpb_last_error()
  std::vector<double> x;          // IterationLimitExceeded(Array1<F>): best x / tau in slack form
  int code = 0;                   // raw lpb status
  std::string to_string() const { return std::string(lpb_strerror(code)) + (message.empty() ? "" : ": " + message); }
};

template <class T>
class Result {
 public:
  Result(T v) : ok_(true), v_(std::move(v)) {}
  Result(LinearProgramError e) : ok_(false), e_(std::move(e)) {}
  bool ok() const { return ok_; }
  const T& value() const { return v_; }
  const LinearProgramError& error() const { return e_; }
  T unwrap() const {
    if (!ok_) throw std::runtime_error(e_.to_string());
    return v_;
  }

 private:
  bool ok_;
  T v_{};
  LinearProgramError e_{};
};

inline LinearProgramError make_error(int code, std::vector<double> x = {}) {
  LinearProgramError e;
  e.code = code;
  e.kind = (code >= 1 && code <= 7) ? static_cast<LinearProgramError::Kind>(code) : LinearProgramError::Device;
  if (code == LPB_ERR_UNSUPPORTED) e.kind = LinearProgramError::InvalidParameter;
  if (code < 0) e.message = lpb_last_error();
  e.x = std::move(x);
  return e;
}

// Row-major dense views of the caller's arrays (ndarray::Array2 / Array1 stand-ins).
struct Array2 {
  const double* data = nullptr;
  int64_t rows = 0, cols = 0;
};
struct Array1 {
  const double* data = nullptr;
  int64_t len = 0;
};

// ------------------------------------------------------------------ linear_program.rs
class Problem;
class ProblemBuilder {
 public:
  explicit ProblemBuilder(Array1 c) : c_(c) {}                 // ProblemBuilder::new  :81-87
  ProblemBuilder& ub(Array2 A, Array1 b) { ub_A_ = A; ub_b_ = b; has_ub_ = true; return *this; }  // :93-96
  ProblemBuilder& eq(Array2 A, Array1 b) { eq_A_ = A; eq_b_ = b; has_eq_ = true; return *this; }  // :102-105
  Result<Problem> build() const;                               // :125-169

 private:
  Array1 c_;
  Array2 ub_A_{}, eq_A_{};
  Array1 ub_b_{}, eq_b_{};
  bool has_ub_ = false, has_eq_ = false;
};

class Problem {
 public:
  static ProblemBuilder target(Array1 c) { return ProblemBuilder(c); }  // :37-39
  const std::vector<double>& A() const { return A_; }                   // row-major m x n
  const std::vector<double>& b() const { return b_; }
  const std::vector<double>& c() const { return c_; }
  int64_t rows() const { return m_; }
  int64_t cols() const { return n_; }
  double c0() const { return c0_; }
  int64_t n_slack() const { return n_slack_; }

 private:
  friend class ProblemBuilder;
  std::vector<double> A_, b_, c_;
  int64_t m_ = 0, n_ = 0, n_slack_ = 0;
  double c0_ = 0.0;
};

inline Result<Problem> ProblemBuilder::build() const {
  const int64_t n_c = c_.len;
  // the reference substitutes (0, n) placeholders for a missing block (:127-130)
  const int64_t rows_ub = has_ub_ ? ub_A_.rows : 0, cols_ub = has_ub_ ? ub_A_.cols : n_c;
  const int64_t rows_eq = has_eq_ ? eq_A_.rows : 0, cols_eq = has_eq_ ? eq_A_.cols : n_c;
  const int64_t len_bub = has_ub_ ? ub_b_.len : 0, len_beq = has_eq_ ? eq_b_.len : 0;
  Problem p;
  int rc = lpb_slack_dims(n_c, rows_ub, cols_ub, len_bub, rows_eq, cols_eq, len_beq, &p.m_, &p.n_, &p.n_slack_);
  if (rc != LPB_OK) return make_error(rc);
  p.A_.assign(static_cast<size_t>(p.m_ * p.n_), 0.0);
  p.b_.assign(static_cast<size_t>(p.m_), 0.0);
  p.c_.assign(static_cast<size_t>(p.n_), 0.0);
  rc = lpb_build_slack_form(c_.data, n_c, ub_A_.data, rows_ub, cols_ub, cols_ub, ub_b_.data, len_bub, eq_A_.data,
                            rows_eq, cols_eq, cols_eq, eq_b_.data, len_beq, p.A_.data(), p.n_, p.b_.data(),
                            p.c_.data());
  if (rc != LPB_OK) return make_error(rc);
  return p;
}

// ------------------------------------------------------------------ solvers/mod.rs
class OptimizeResult {
 public:
  OptimizeResult() = default;
  OptimizeResult(std::vector<double> x, double fun, int64_t iteration)
      : x_(std::move(x)), fun_(fun), iteration_(iteration) {}
  int64_t iteration() const { return iteration_; }   // :36-38
  const double& fun() const { return fun_; }         // :41-43
  const std::vector<double>& x() const { return x_; }  // :46-48

 private:
  std::vector<double> x_;
  double fun_ = 0.0;
  int64_t iteration_ = 0;
};

class Solver {  // solvers/mod.rs:12-16
 public:
  virtual ~Solver() = default;
  virtual Result<OptimizeResult> solve(const Problem& problem) const = 0;
};

enum class EquationSolverType { Cholesky = LPB_SOLVER_CHOLESKY, Inverse = LPB_SOLVER_INVERSE,
                                LeastSquares = LPB_SOLVER_LEAST_SQUARES };

// The phase calls of the C ABI as the `Dev` that ipm_driver.hpp drives.
struct CabiDevice {
  lpb_ctx* ctx;
  int blind_start() { return lpb_blind_start(ctx); }
  int residuals(double tau, double kappa, lpb_residual_scalars* o) { return lpb_residuals(ctx, tau, kappa, o); }
  int form_and_factor() { return lpb_form_and_factor(ctx); }
  int direction(const lpb_direction_in& in, double tau, double kappa, lpb_direction_out* o) {
    return lpb_direction(ctx, &in, tau, kappa, o);
  }
  int assemble_delta(double d_tau, double axz[2]) { return lpb_assemble_delta(ctx, d_tau, axz); }
  int do_step(double alpha, int ip) { return lpb_do_step(ctx, alpha, ip); }
};

class InteriorPoint;
class InteriorPointBuilder {  // interior_point/mod.rs:41-138
 public:
  InteriorPointBuilder() { lpb_options_default(&o_); }
  InteriorPointBuilder& tol(double v) { o_.tol = v; return *this; }
  InteriorPointBuilder& disp(bool v) { o_.disp = v; return *this; }
  InteriorPointBuilder& ip(bool v) { o_.ip = v; return *this; }
  InteriorPointBuilder& solver_type(EquationSolverType t) { o_.solver_type = static_cast<int>(t); return *this; }
  InteriorPointBuilder& alpha0(double v) { o_.alpha0 = v; return *this; }
  InteriorPointBuilder& max_iter(int64_t v) { o_.max_iter = v; return *this; }
  Result<InteriorPoint> build() const;

 private:
  lpb_options o_;
};

class InteriorPoint : public Solver {  // interior_point/mod.rs:145-240
 public:
  InteriorPoint() { lpb_options_default(&o_); }
  explicit InteriorPoint(const lpb_options& o) : o_(o) {}
  static InteriorPoint default_() { return InteriorPoint(); }          // Default::default  :154-159
  static InteriorPointBuilder custom() { return InteriorPointBuilder(); }  // :195-197
  bool operator==(const InteriorPoint& r) const {
    return o_.tol == r.o_.tol && o_.disp == r.o_.disp && o_.ip == r.o_.ip && o_.solver_type == r.o_.solver_type &&
           o_.alpha0 == r.o_.alpha0 && o_.max_iter == r.o_.max_iter;
  }

  // Solver::solve (:161-169): solve_normal_form on the GPU, then denormalize.
  Result<OptimizeResult> solve(const Problem& problem) const override {
    int rc = lpb_options_validate(&o_);
    if (rc != LPB_OK) return make_error(rc);
    lpb_ctx* ctx = nullptr;
    rc = lpb_create(&ctx, problem.rows(), problem.cols(), problem.A().data(), problem.cols(), problem.b().data(),
                    problem.c().data(), problem.c0(), LPB_MEM_HOST, nullptr);
    if (rc != LPB_OK) return make_error(rc);
    CabiDevice dev{ctx};
    lpb::SolveOutput out;
    rc = lpb::solve_normal_form(dev, o_, problem.cols(), problem.c0(), &out);  // mod.rs:199-240, host-driven
    std::vector<double> x_slack(static_cast<size_t>(problem.cols()));
    double fun = 0.0;
    if (rc == LPB_OK || rc == LPB_ERR_ITERATION_LIMIT_EXCEEDED) {
      const int rc2 = lpb_extract_x(ctx, out.tau, x_slack.data(), &fun);  // x / tau, c.x + c0
      if (rc2 != LPB_OK) rc = rc2;
    }
    lpb_destroy(ctx);
    if (rc == LPB_ERR_ITERATION_LIMIT_EXCEEDED) return make_error(rc, std::move(x_slack));  // mod.rs:237-239
    if (rc != LPB_OK) return make_error(rc);
    x_slack.resize(static_cast<size_t>(problem.cols() - problem.n_slack()));  // linear_program.rs:65-69
    return OptimizeResult(std::move(x_slack), fun, out.iterations);
  }

 private:
  lpb_options o_;
};

inline Result<InteriorPoint> InteriorPointBuilder::build() const {
  if (o_.alpha0 <= 0.0 || o_.alpha0 >= 1.0) {  // mod.rs:119-123
    LinearProgramError e = make_error(LPB_ERR_INVALID_PARAMETER);
    e.message = "Alpha0 must be between 0 and 1 (exclusive)";
    return e;
  }
  if (o_.tol <= 0.0) {  // mod.rs:124-128
    LinearProgramError e = make_error(LPB_ERR_INVALID_PARAMETER);
    e.message = "The tolerance must be nonnegative.";
    return e;
  }
  return InteriorPoint(o_);
}

}  // namespace ripped
