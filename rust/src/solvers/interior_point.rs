//! The MOSEK-style homogeneous predictor-corrector interior-point solver, B200 edition.
//!
//! Public surface = `/root/reference/src/solvers/interior_point/mod.rs`: `InteriorPointBuilder` (`:41-138`, same six
//! options, same defaults, same validation), `InteriorPoint::{default, custom}` (`:154-159`, `:195-197`),
//! `impl Solver for InteriorPoint` (`:161-169`), `EquationSolverType` (`newton_equations.rs:36-46`).
//!
//! `solve_normal_form` below is the reference's loop (`mod.rs:199-240`) with every array expression replaced by one
//! phase call of `include/lpb200.h`; the scalar logic (`feasible_point.rs:53-72, 110-165`, `rhat.rs`, `delta.rs:29-38`,
//! `residual.rs:13-44`, `indicators.rs:37-83`) stays here on the host, operation for operation.  The C++ twin of this
//! file is `lp_b200/csrc/ipm_driver.hpp` (that one is compiled and tested in this repository; this one is UNBUILT).
#![allow(unsafe_code)] // phase calls through ffi

use std::fmt;
use std::os::raw::c_int;

use ndarray::Array1;

use super::{OptimizeResult, Solver};
use crate::error::LinearProgramError;
use crate::ffi::{self, CtxGuard};
use crate::float::Float;
use crate::linear_program::Problem;

type LpResult<T> = Result<T, LinearProgramError<f64>>;

/// Which factorisation solves the normal equations (reference `newton_equations.rs:36-46`).  Only `Cholesky` runs
/// on the B200 path; the other two are the reference's CPU fallback chain and are rejected with
/// `InvalidParameter` at `solve` time.
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub enum EquationSolverType {
    /// Blocked Cholesky of `A D A^T` on the FP64 tensor cores.
    Cholesky,
    /// Not available on the B200 path.
    Inverse,
    /// Not available on the B200 path.
    LeastSquares,
}

impl EquationSolverType {
    fn code(self) -> i32 {
        match self {
            EquationSolverType::Cholesky => 0,
            EquationSolverType::Inverse => 1,
            EquationSolverType::LeastSquares => 2,
        }
    }
}

/// Builder of [`InteriorPoint`]: start from the defaults, change what you need, `build()` validates.
pub struct InteriorPointBuilder<F> {
    tol: F,
    disp: bool,
    ip: bool,
    solver_type: EquationSolverType,
    alpha0: F,
    max_iter: usize,
}

impl<F: Float> InteriorPointBuilder<F> {
    pub(crate) fn new() -> Self {
        // reference mod.rs:51-60
        InteriorPointBuilder {
            tol: F::from_f64(1e-8),
            disp: false,
            ip: true,
            solver_type: EquationSolverType::Cholesky,
            alpha0: F::from_f64(0.99995),
            max_iter: 1000,
        }
    }

    /// Convergence tolerance on the indicators (small, positive).
    pub fn tol(mut self, tol: F) -> Self {
        self.tol = tol;
        self
    }

    /// Print the indicators at every iteration (same columns as the reference).
    pub fn disp(mut self, disp: bool) -> Self {
        self.disp = disp;
        self
    }

    /// Use the "improved" initial point handling of the first iteration.
    pub fn ip(mut self, ip: bool) -> Self {
        self.ip = ip;
        self
    }

    /// Equation solver to use; only `Cholesky` exists on the B200 path.
    pub fn solver_type(mut self, solver_type: EquationSolverType) -> Self {
        self.solver_type = solver_type;
        self
    }

    /// Step-size multiplier, `0 < alpha0 < 1`.
    pub fn alpha0(mut self, alpha0: F) -> Self {
        self.alpha0 = alpha0;
        self
    }

    /// Iteration limit.
    pub fn max_iter(mut self, max_iter: usize) -> Self {
        self.max_iter = max_iter;
        self
    }

    /// Validate (reference `mod.rs:118-128`) and create the solver.
    pub fn build(self) -> Result<InteriorPoint<F>, LinearProgramError<F>> {
        if self.alpha0.to_f64() <= 0.0 || self.alpha0.to_f64() >= 1.0 {
            return Err(LinearProgramError::InvalidParameter("Alpha0 must be between 0 and 1 (exclusive)"));
        }
        if self.tol.to_f64() <= 0.0 {
            return Err(LinearProgramError::InvalidParameter("The tolerance must be nonnegative."));
        }
        Ok(InteriorPoint {
            tol: self.tol,
            disp: self.disp,
            ip: self.ip,
            solver_type: self.solver_type,
            alpha0: self.alpha0,
            max_iter: self.max_iter,
        })
    }
}

/// The interior-point solver; `InteriorPoint::default()` or `InteriorPoint::custom()...build()`.
#[derive(PartialEq, Debug)]
pub struct InteriorPoint<F> {
    tol: F,
    disp: bool,
    ip: bool,
    solver_type: EquationSolverType,
    alpha0: F,
    max_iter: usize,
}

impl<F: Float> Default for InteriorPoint<F> {
    fn default() -> Self {
        InteriorPointBuilder::new().build().unwrap() // the defaults pass validation (reference mod.rs:157)
    }
}

impl<F: Float> Solver<F> for InteriorPoint<F> {
    /// Reference `mod.rs:161-169`: run the loop, then `fun = c . x_slack + c0` and strip the slack variables.
    /// `fun` is formed on the device by `lpb_extract_x` (same dot product, `linear_program.rs:61-63`).
    /// `F = f32`: the slack form is widened to FP64 for the upload and `x` / `fun` are narrowed back.
    fn solve(&self, problem: &Problem<F>) -> Result<OptimizeResult<F>, LinearProgramError<F>> {
        let a = problem.A().mapv(Float::to_f64);
        let b = problem.b().mapv(Float::to_f64);
        let c = problem.c().mapv(Float::to_f64);
        match self.solve_normal_form(&a, &b, &c, problem.c0().to_f64()) {
            Ok((x_slack, fun, iteration)) => {
                let x = problem.denormalize_x_into(x_slack.mapv(F::from_f64));
                Ok(OptimizeResult::new(x, F::from_f64(fun), iteration))
            }
            Err(e) => Err(narrow_error(e)),
        }
    }
}

/// The loop reports in FP64; the caller's error type carries `F`.
fn narrow_error<F: Float>(e: LinearProgramError<f64>) -> LinearProgramError<F> {
    match e {
        LinearProgramError::Unconstrained => LinearProgramError::Unconstrained,
        LinearProgramError::NumericalProblem => LinearProgramError::NumericalProblem,
        LinearProgramError::InvalidParameter(s) => LinearProgramError::InvalidParameter(s),
        LinearProgramError::IncompatibleInputDimensions => LinearProgramError::IncompatibleInputDimensions,
        LinearProgramError::Infeasible => LinearProgramError::Infeasible,
        LinearProgramError::Unbounded => LinearProgramError::Unbounded,
        LinearProgramError::IterationLimitExceeded(x) => LinearProgramError::IterationLimitExceeded(x.mapv(F::from_f64)),
    }
}

// ------------------------------------------------------------------------------------------------ scalar pieces

/// `residual.rs:5-10` evaluated at the blind start: the denominators of the indicators.
#[derive(Clone, Copy, Debug)]
struct InitialResiduals {
    rho_p: f64,
    rho_d: f64,
    rho_g: f64,
    rho_mu: f64,
}

/// `residual.rs:13-44` from the reduction scalars of one residual sweep.
fn residual_values(rs: &ffi::lpb_residual_scalars, tau: f64, kappa: f64, n: usize) -> InitialResiduals {
    InitialResiduals {
        rho_p: rs.nrm_rp,
        rho_d: rs.nrm_rd,
        rho_g: (kappa + rs.cx - rs.by).abs(),
        rho_mu: (rs.xz + tau * kappa) / (n + 1) as f64,
    }
}

/// `indicators.rs:8-23`.
#[derive(Clone, Copy, Debug)]
struct Indicators {
    rho_p: f64,
    rho_d: f64,
    rho_A: f64,
    rho_g: f64,
    rho_mu: f64,
    obj: f64,
    bty: f64,
}

/// `indicators.rs:85-90`.
#[derive(Clone, Copy, PartialEq, Eq, Debug)]
enum Status {
    Optimal,
    Infeasible,
    Unbounded,
    Unfinished,
}

impl Indicators {
    /// `indicators.rs:37-55`.
    fn from_scalars(rs: &ffi::lpb_residual_scalars, ini: &InitialResiduals, tau: f64, kappa: f64, n: usize,
                    c0: f64) -> Self {
        let now = residual_values(rs, tau, kappa, n);
        Indicators {
            obj: rs.cx / tau + c0,
            bty: rs.by,
            rho_A: (rs.cx - rs.by).abs() / (tau + rs.by.abs()),
            rho_p: now.rho_p / ini.rho_p.max(1.0),
            rho_d: now.rho_d / ini.rho_d.max(1.0),
            rho_g: now.rho_g / ini.rho_g.max(1.0),
            rho_mu: now.rho_mu / ini.rho_mu,
        }
    }

    /// `indicators.rs:66-83`: the infeasibility test has priority over the optimality test; every comparison is a
    /// strict `<` / `>`, so NaN indicators give `Unfinished`.
    fn status(&self, tau: f64, kappa: f64, tol: f64) -> Status {
        let tau_too_small = tau < tol * kappa.max(1.0);
        let inf1 = (self.rho_p < tol && self.rho_d < tol && self.rho_g < tol) && tau_too_small;
        let inf2 = self.rho_mu < tol && tau_too_small;
        if inf1 || inf2 {
            return if self.bty > tol { Status::Infeasible } else { Status::Unbounded };
        }
        if self.rho_p < tol && self.rho_d < tol && self.rho_A < tol {
            return Status::Optimal;
        }
        Status::Unfinished
    }
}

impl fmt::Display for Indicators {
    /// `indicators.rs:25-33`.
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        write!(f, "{:3.8}\t{:3.8}\t{:3.8}\t{:3.8}\t{:8.3}", self.rho_p, self.rho_d, self.rho_g, self.rho_mu, self.obj)
    }
}

/// One `Delta::compute` worth of host scalars (`delta.rs:29-32`, `:38`).
fn delta_scalars(g_hat: f64, tk_hat: f64, tau: f64, kappa: f64, d: &ffi::lpb_direction_out) -> (f64, f64) {
    let d_tau = (g_hat + 1.0 / tau * tk_hat - (-d.cu + d.bv)) / (1.0 / tau * kappa + (-d.cp + d.bq));
    let d_kappa = 1.0 / tau * (tk_hat - kappa * d_tau);
    (d_tau, d_kappa)
}

/// `feasible_point.rs:53-72`; `axz` = (alpha_x, alpha_z) from the device ratio test, already min'ed with 1;
/// `alpha0` multiplies AFTER the min with 1.
fn step_size(axz: [f64; 2], tau: f64, d_tau: f64, kappa: f64, d_kappa: f64, alpha0: f64) -> f64 {
    let alpha_tau = if d_tau < 0.0 { 1.0f64.min(tau / -d_tau) } else { 1.0 };
    let alpha_kappa = if d_kappa < 0.0 { 1.0f64.min(kappa / -d_kappa) } else { 1.0 };
    1.0f64.min(axz[0]).min(alpha_tau).min(axz[1]).min(alpha_kappa) * alpha0
}

/// `feasible_point.rs:155-165`.
fn update_gamma(ip: bool, alpha: f64) -> f64 {
    if ip {
        return 10.0;
    }
    let beta1 = 0.1f64;
    (1.0 - alpha).powi(2) * beta1.min(1.0 - alpha)
}

/// Non-zero status of a phase call -> error (codes 1..6 and the negative device codes; 7 never comes from a phase).
fn check(rc: c_int) -> LpResult<()> {
    if rc == ffi::LPB_OK {
        return Ok(());
    }
    Err(device_error(rc))
}

fn device_error(rc: c_int) -> LinearProgramError<f64> {
    if rc < 0 {
        eprintln!("liblpb200: status {} -- {}", rc, ffi::last_error());
    }
    LinearProgramError::from_code(rc)
}

/// `x / tau` in slack form and `c . (x / tau) + c0` (`mod.rs:231`, `:237-239`; `linear_program.rs:61-63`).
fn extract_x(ctx: &CtxGuard, tau: f64, n: usize) -> LpResult<(Array1<f64>, f64)> {
    let mut x = vec![0.0f64; n];
    let mut fun = 0.0f64;
    // SAFETY: x has n elements, the context holds an n-column problem; fun is a valid out-pointer.
    check(unsafe { ffi::lpb_extract_x(ctx.raw(), tau, x.as_mut_ptr(), &mut fun) })?;
    Ok((Array1::from(x), fun))
}

impl<F: Float> InteriorPoint<F> {
    /// Customise the solver through the builder (reference `mod.rs:195-197`).
    pub fn custom() -> InteriorPointBuilder<F> {
        InteriorPointBuilder::new()
    }

    fn options(&self) -> ffi::lpb_options {
        ffi::lpb_options {
            tol: self.tol.to_f64(),
            disp: self.disp as i32,
            ip: self.ip as i32,
            solver_type: self.solver_type.code(),
            reserved: 0,
            alpha0: self.alpha0.to_f64(),
            max_iter: self.max_iter as i64,
        }
    }

    /// The loop of reference `mod.rs:199-240`.  Returns `(x / tau in slack form, c . x / tau + c0, iterations)`.
    fn solve_normal_form(&self, a: &ndarray::Array2<f64>, b: &Array1<f64>, c: &Array1<f64>, c0: f64)
                         -> LpResult<(Array1<f64>, f64, usize)> {
        let opts = self.options();
        let (tol, alpha0) = (opts.tol, opts.alpha0);
        // SAFETY: opts is a valid, initialised struct.
        check(unsafe { ffi::lpb_options_validate(&opts) })?; // Inverse / LeastSquares -> InvalidParameter

        let (m, n) = a.dim();
        let a = a.as_standard_layout(); // row-major, as build() made it (no copy)
        let b = b.as_standard_layout();
        let c = c.as_standard_layout();
        let ctx = CtxGuard::create(
            m,
            n,
            a.as_slice().expect("standard layout"),
            b.as_slice().expect("standard layout"),
            c.as_slice().expect("standard layout"),
            c0,
        )
        .map_err(device_error)?;
        let h = ctx.raw();

        let (mut tau, mut kappa) = (1.0f64, 1.0f64); // feasible_point.rs:29-30
        // SAFETY (all phase calls below): `h` is the live context owned by `ctx`; out-pointers are valid locals.
        check(unsafe { ffi::lpb_blind_start(h) })?;
        let mut rs = ffi::lpb_residual_scalars::default();
        check(unsafe { ffi::lpb_residuals(h, tau, kappa, &mut rs) })?;
        let ini = residual_values(&rs, tau, kappa, n); // feasible_point.rs:32
        let mut indicators = Indicators::from_scalars(&rs, &ini, tau, kappa, n, c0); // mod.rs:206
        if self.disp {
            // mod.rs:208-211
            println!("alpha     \trho_p     \trho_d     \trho_g     \trho_mu    \tobj       ");
            println!("1.00000000\t{}", indicators);
        }

        let mut ip = self.ip;
        for iteration in 1..=self.max_iter {
            // ---- get_delta (feasible_point.rs:110-152)
            let mut gamma = if ip { 1.0 } else { 0.0 }; // :119
            let mut eta = if ip { 1.0 } else { 1.0 - gamma }; // :120
            let r_g = rs.cx - rs.by + kappa; // :124
            let mu = (rs.xz + tau * kappa) / (n + 1) as f64; // :125

            match unsafe { ffi::lpb_form_and_factor(h) } {
                // newton_equations.rs:48-64; a failed factorisation is final (:63), no fallback chain
                ffi::LPB_OK => {}
                ffi::LPB_ERR_NUMERICAL_PROBLEM => return Err(LinearProgramError::NumericalProblem),
                rc => return Err(device_error(rc)),
            }

            // predictor (rhat.rs:17-35)
            let mut din = ffi::lpb_direction_in { corrector: 0, ip: ip as i32, eta, gamma, mu, alpha: 0.0 };
            let mut dout = ffi::lpb_direction_out::default();
            check(unsafe { ffi::lpb_direction(h, &din, tau, kappa, &mut dout) })?;
            if dout.nan_pq != 0 {
                return Err(LinearProgramError::NumericalProblem); // newton_equations.rs:190-194
            }
            let (mut d_tau, mut d_kappa) = delta_scalars(r_g * eta, gamma * mu - tau * kappa, tau, kappa, &dout);
            let mut axz = [1.0f64; 2];
            check(unsafe { ffi::lpb_assemble_delta(h, d_tau, axz.as_mut_ptr()) })?; // delta.rs:33-37 + ratio test

            let alpha = step_size(axz, tau, d_tau, kappa, d_kappa, 1.0); // feasible_point.rs:134
            gamma = update_gamma(ip, alpha); // :135
            eta = if ip { 1.0 } else { 1.0 - gamma }; // :136

            // corrector (rhat.rs:37-75)
            let tk_hat = if ip {
                (1.0 - alpha) * gamma * mu - tau * kappa - alpha * alpha * d_tau * d_kappa // :51-60
            } else {
                gamma * mu - tau * kappa - d_tau * d_kappa // :62-66
            };
            din = ffi::lpb_direction_in { corrector: 1, ip: ip as i32, eta, gamma, mu, alpha };
            check(unsafe { ffi::lpb_direction(h, &din, tau, kappa, &mut dout) })?;
            if dout.nan_pq != 0 {
                return Err(LinearProgramError::NumericalProblem);
            }
            let (dt, dk) = delta_scalars(r_g * eta, tk_hat, tau, kappa, &dout);
            d_tau = dt;
            d_kappa = dk;
            check(unsafe { ffi::lpb_assemble_delta(h, d_tau, axz.as_mut_ptr()) })?;

            // ---- step (mod.rs:216-223)
            let alpha = if ip { 1.0 } else { step_size(axz, tau, d_tau, kappa, d_kappa, alpha0) };
            check(unsafe { ffi::lpb_do_step(h, alpha, ip as c_int) })?; // feasible_point.rs:76-106
            tau += d_tau * alpha;
            kappa += d_kappa * alpha;
            if ip {
                tau = tau.max(1.0); // :92-93
                kappa = kappa.max(1.0);
            }
            ip = false;

            // ---- indicators (mod.rs:225-235)
            check(unsafe { ffi::lpb_residuals(h, tau, kappa, &mut rs) })?;
            indicators = Indicators::from_scalars(&rs, &ini, tau, kappa, n, c0);
            if self.disp {
                println!("{:3.8}\t{}", alpha, indicators); // mod.rs:227-229
            }
            match indicators.status(tau, kappa, tol) {
                Status::Optimal => {
                    let (x, fun) = extract_x(&ctx, tau, n)?;
                    return Ok((x, fun, iteration)); // mod.rs:231
                }
                Status::Infeasible => return Err(LinearProgramError::Infeasible),
                Status::Unbounded => return Err(LinearProgramError::Unbounded),
                Status::Unfinished => {}
            }
        }
        let (x, _) = extract_x(&ctx, tau, n)?;
        Err(LinearProgramError::IterationLimitExceeded(x)) // mod.rs:237-239: best x / tau in slack form
    }
}

#[cfg(test)]
mod tests {
    //! The reference's own unit tests (`mod.rs:243-345`, `lib.rs:77-114`), unchanged in what they assert.
    use super::*;
    use crate::prelude::*;
    use approx::assert_abs_diff_eq;
    use ndarray::array;

    #[test]
    fn default_builder_doesnt_panic() {
        assert_eq!(InteriorPoint::<f64>::default(), InteriorPoint::custom().build().unwrap());
    }

    #[test]
    fn test_interior_point_builder() {
        let A_ub = array![[-3f64, 1.], [1., 2.]];
        let b_ub = array![6., 4.];
        let A_eq = array![[1., 1.]];
        let b_eq = array![1.];
        let c = array![-1., 4.];
        let problem = Problem::target(&c).ub(&A_ub, &b_ub).eq(&A_eq, &b_eq).build().unwrap();
        let res = InteriorPoint::default().solve(&problem).unwrap();
        assert_abs_diff_eq!(*res.x(), array![1., 0.], epsilon = 1e-6);
    }

    #[test]
    fn fallback_solver_types_are_rejected_not_silently_mapped() {
        let A_ub = array![[-3f64, 1.], [1., 2.]];
        let b_ub = array![6., 4.];
        let c = array![-1., 4.];
        let problem = Problem::target(&c).ub(&A_ub, &b_ub).build().unwrap();
        for st in [EquationSolverType::Inverse, EquationSolverType::LeastSquares] {
            let solver = InteriorPoint::custom().solver_type(st).build().unwrap();
            assert!(matches!(solver.solve(&problem), Err(LinearProgramError::InvalidParameter(_))));
        }
    }

    #[test]
    fn test_linprog_eq_only() {
        let A_eq = array![[2.0, 1.0, 0.0], [0.0, 2.0, 1.0], [1.0, 0.0, 2.0]];
        let b_eq = array![1.0, 2.0, 3.0];
        let c = array![-1.0, 4.0, -1.2];
        let problem = Problem::target(&c).eq(&A_eq, &b_eq).build().unwrap();
        let res = InteriorPoint::default().solve(&problem).unwrap();
        assert_abs_diff_eq!(*res.x(), array![1. / 3., 1. / 3., 4. / 3.], epsilon = 1e-6);
    }

    #[test]
    fn test_linprog_ub_only() {
        let A_ub = array![[2.0, 1.0, 0.0], [0.0, 2.0, 1.0], [1.0, 0.0, 2.0]];
        let b_ub = array![1.0, 2.0, 3.0];
        let c = array![-1.0, 4.0, -1.2];
        let problem = Problem::target(&c).ub(&A_ub, &b_ub).build().unwrap();
        let res = InteriorPoint::default().solve(&problem).unwrap();
        assert_abs_diff_eq!(*res.x(), array![0.5, 0.0, 1.25], epsilon = 1e-6);
    }

    #[test]
    fn f32_problems_are_solved_in_f64_and_narrowed() {
        let A_ub = array![[-3f32, 1.], [1., 2.]];
        let b_ub = array![6f32, 4.];
        let c = array![-1f32, 4.];
        let problem = Problem::target(&c).ub(&A_ub, &b_ub).build().unwrap();
        let res = InteriorPoint::<f32>::default().solve(&problem).unwrap();
        assert_abs_diff_eq!(*res.x(), array![4f32, 0.], epsilon = 1e-5);
    }

    #[test]
    fn iteration_limit_carries_slack_form_x() {
        let A_ub = array![[-3f64, 1.], [1., 2.]];
        let b_ub = array![6., 4.];
        let c = array![-1., 4.];
        let problem = Problem::target(&c).ub(&A_ub, &b_ub).build().unwrap();
        let solver = InteriorPoint::custom().max_iter(1).build().unwrap();
        match solver.solve(&problem) {
            Err(LinearProgramError::IterationLimitExceeded(x)) => assert_eq!(x.len(), 4), // 2 variables + 2 slacks
            _ => panic!("expected IterationLimitExceeded"),
        }
    }
}
