// test_ripped.cpp -- the reference's own unit tests and doctests, restated against the C++ host mirror
// (lp_b200/host/ripped.hpp) over the C ABI.  Paths are relative to /root/reference/src.
//   host-only cases (always run):   builder validation, slack form, error variants
//   GPU cases (run with --gpu):     lib.rs:84-113, interior_point/mod.rs:181-192, 256-273, 319-344,
//                                   examples/symmetric.rs:10-25 (N = 1000, x == 1 to 1e-10)
// Exit code 0 = all assertions held.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../lp_b200/host/ripped.hpp"

using namespace ripped;

static int failures = 0;
#define CHECK(cond)                                                        \
  do {                                                                     \
    if (!(cond)) {                                                         \
      std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond);          \
      ++failures;                                                          \
    }                                                                      \
  } while (0)

static bool close_to(const std::vector<double>& x, const std::vector<double>& want, double eps) {
  if (x.size() != want.size()) return false;
  for (size_t i = 0; i < x.size(); ++i)
    if (!(std::fabs(x[i] - want[i]) <= eps)) return false;
  return true;
}

static Array2 mat(const std::vector<double>& v, int64_t r, int64_t c) { return Array2{v.data(), r, c}; }
static Array1 vec(const std::vector<double>& v) { return Array1{v.data(), (int64_t)v.size()}; }

static void host_only_cases() {
  // interior_point/mod.rs:250-254 default_builder_doesnt_panic
  CHECK(InteriorPoint::default_() == InteriorPoint::custom().build().unwrap());
  // mod.rs:118-128 validation
  CHECK(!InteriorPoint::custom().alpha0(1.0).build().ok());
  CHECK(InteriorPoint::custom().alpha0(1.0).build().error().kind == LinearProgramError::InvalidParameter);
  CHECK(!InteriorPoint::custom().alpha0(0.0).build().ok());
  CHECK(!InteriorPoint::custom().tol(0.0).build().ok());
  CHECK(InteriorPoint::custom().tol(1e-6).build().ok());
  // linear_program.rs:134-143
  std::vector<double> c = {1.0, 2.0};
  CHECK(Problem::target(vec(c)).build().error().kind == LinearProgramError::Unconstrained);
  std::vector<double> A3 = {1.0, 2.0, 3.0}, b1 = {1.0}, b2 = {1.0, 2.0}, A2 = {1.0, 2.0};
  CHECK(Problem::target(vec(c)).ub(mat(A3, 1, 3), vec(b1)).build().error().kind ==
        LinearProgramError::IncompatibleInputDimensions);
  CHECK(Problem::target(vec(c)).ub(mat(A2, 1, 2), vec(b2)).build().error().kind ==
        LinearProgramError::IncompatibleInputDimensions);
  // lib.rs:77-96 test_problem_interface: slack form [[A_ub, I], [A_eq, 0]]
  std::vector<double> cc = {-1.0, 4.0}, A_ub = {-3.0, 1.0, 1.0, 2.0}, b_ub = {6.0, 4.0}, A_eq = {1.0, 1.0}, b_eq = {1.0};
  auto p = Problem::target(vec(cc)).ub(mat(A_ub, 2, 2), vec(b_ub)).eq(mat(A_eq, 1, 2), vec(b_eq)).build();
  CHECK(p.ok());
  const std::vector<double> wantA = {-3, 1, 1, 0, 1, 2, 0, 1, 1, 1, 0, 0};
  CHECK(p.value().A() == wantA);
  CHECK((p.value().b() == std::vector<double>{6, 4, 1}));
  CHECK((p.value().c() == std::vector<double>{-1, 4, 0, 0}));
  CHECK(p.value().n_slack() == 2 && p.value().rows() == 3 && p.value().cols() == 4);
}

static void gpu_cases() {
  std::vector<double> c = {-1.0, 4.0}, A_ub = {-3.0, 1.0, 1.0, 2.0}, b_ub = {6.0, 4.0}, A_eq = {1.0, 1.0}, b_eq = {1.0};
  {  // lib.rs:84-113 test_interior_point_interface / mod.rs:256-273 test_interior_point_builder
    auto problem = Problem::target(vec(c)).ub(mat(A_ub, 2, 2), vec(b_ub)).eq(mat(A_eq, 1, 2), vec(b_eq)).build().unwrap();
    auto solver = InteriorPoint::custom().tol(1e-8).disp(false).ip(true)
                      .solver_type(EquationSolverType::Cholesky).alpha0(0.99995).max_iter(1000).build().unwrap();
    auto res = solver.solve(problem);
    CHECK(res.ok());
    CHECK(close_to(res.value().x(), {1.0, 0.0}, 1e-6));
    CHECK(std::fabs(res.value().fun() + 1.0) < 1e-6);
    CHECK(res.value().iteration() >= 1);
    // the Solver trait object works like `&dyn Solver`
    const Solver& dyn = solver;
    CHECK(close_to(dyn.solve(problem).unwrap().x(), {1.0, 0.0}, 1e-6));
    // mod.rs:275-317: Inverse / LeastSquares are NOT silently mapped on the GPU path
    auto inv = InteriorPoint::custom().solver_type(EquationSolverType::Inverse).build().unwrap().solve(problem);
    CHECK(!inv.ok() && inv.error().kind == LinearProgramError::InvalidParameter);
    // mod.rs:237-239: IterationLimitExceeded carries x / tau in slack form
    auto lim = InteriorPoint::custom().max_iter(1).build().unwrap().solve(problem);
    CHECK(!lim.ok() && lim.error().kind == LinearProgramError::IterationLimitExceeded && lim.error().x.size() == 4);
  }
  {  // mod.rs:181-192 doctest of InteriorPoint::custom: ub only, x = [4, 0]
    auto problem = Problem::target(vec(c)).ub(mat(A_ub, 2, 2), vec(b_ub)).build().unwrap();
    CHECK(close_to(InteriorPoint::default_().solve(problem).unwrap().x(), {4.0, 0.0}, 1e-6));
  }
  std::vector<double> c3 = {-1.0, 4.0, -1.2}, A3 = {2, 1, 0, 0, 2, 1, 1, 0, 2}, b3 = {1.0, 2.0, 3.0};
  {  // mod.rs:319-331 test_linprog_eq_only
    auto problem = Problem::target(vec(c3)).eq(mat(A3, 3, 3), vec(b3)).build().unwrap();
    CHECK(close_to(InteriorPoint::default_().solve(problem).unwrap().x(), {1.0 / 3.0, 1.0 / 3.0, 4.0 / 3.0}, 1e-6));
  }
  {  // mod.rs:332-344 test_linprog_ub_only
    auto problem = Problem::target(vec(c3)).ub(mat(A3, 3, 3), vec(b3)).build().unwrap();
    CHECK(close_to(InteriorPoint::default_().solve(problem).unwrap().x(), {0.5, 0.0, 1.25}, 1e-6));
  }
  {  // examples/symmetric.rs:10-25
    const int N = 1000;
    std::vector<double> A((size_t)N * N, 1.0), b((size_t)N, (double)(N - 1)), cn((size_t)N, -1.0);
    for (int i = 0; i < N; ++i) A[(size_t)i * N + i] = 0.0;
    auto problem = Problem::target(vec(cn)).ub(mat(A, N, N), vec(b)).build().unwrap();
    auto res = InteriorPoint::default_().solve(problem).unwrap();
    CHECK(close_to(res.x(), std::vector<double>((size_t)N, 1.0), 1e-10));
    std::printf("symmetric: fun=%.9f iterations=%lld\n", res.fun(), (long long)res.iteration());
  }
  {  // Infeasible / Unbounded variants (untested by the reference; statuses per indicators.rs:66-83)
    std::vector<double> c1 = {1.0, 1.0}, A1 = {1.0, 1.0}, bneg = {-1.0};
    auto inf = InteriorPoint::default_().solve(Problem::target(vec(c1)).ub(mat(A1, 1, 2), vec(bneg)).build().unwrap());
    CHECK(!inf.ok() && inf.error().kind == LinearProgramError::Infeasible);
    std::vector<double> c2 = {-1.0, 0.0}, A2 = {1.0, -1.0}, bpos = {1.0};
    auto unb = InteriorPoint::default_().solve(Problem::target(vec(c2)).ub(mat(A2, 1, 2), vec(bpos)).build().unwrap());
    CHECK(!unb.ok() && unb.error().kind == LinearProgramError::Unbounded);
  }
}

int main(int argc, char** argv) {
  const bool gpu = argc > 1 && std::strcmp(argv[1], "--gpu") == 0;
  host_only_cases();
  if (gpu) {
    if (lpb_device_count() <= 0) {
      std::printf("--gpu requested but no CUDA device\n");
      return 2;
    }
    gpu_cases();
  } else {
    // without a GPU the compute path must fail loudly (no CPU fallback)
    if (lpb_device_count() <= 0) {
      std::vector<double> c = {-1.0, 4.0}, A_ub = {-3.0, 1.0, 1.0, 2.0}, b_ub = {6.0, 4.0};
      auto problem = Problem::target(vec(c)).ub(mat(A_ub, 2, 2), vec(b_ub)).build().unwrap();
      auto res = InteriorPoint::default_().solve(problem);
      CHECK(!res.ok() && res.error().code == LPB_ERR_NO_DEVICE);
    }
  }
  std::printf(failures ? "%d FAILURES\n" : "ok (%d failures)\n", failures);
  return failures ? 1 : 0;
}
