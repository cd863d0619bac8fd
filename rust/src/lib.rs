//! UNBUILT in this image (no Rust toolchain): the Rust side of the drop-in.
//!
//! Same public surface as the reference crate (`/root/reference/src/lib.rs:71-75`,
//! `prelude.rs:3-11`): `Problem::target(&c).ub(..).eq(..).build()`,
//! `InteriorPoint::default().solve(&problem)`, `OptimizeResult`, `LinearProgramError`.
//! `Problem` / `ProblemBuilder` / `LinearProgramError` / `OptimizeResult` are the reference's own
//! (pure host code, unchanged); only `impl Solver for InteriorPoint` changes: instead of calling
//! `solve_normal_form` on ndarray it drives the SAME loop through the phase calls of
//! `include/lpb200.h` (the host keeps tau, kappa and all control flow; the GPU keeps A, M, x, y, z).
#![allow(non_snake_case)]

pub mod ffi {
    //! `extern "C"` declarations of include/lpb200.h.  The reference denies `unsafe_code`
    //! (`.cargo/config.toml:6`); this module is the one place that has to relax it.
    #![allow(unsafe_code, non_camel_case_types)]
    use std::os::raw::{c_char, c_int, c_void};

    #[repr(C)]
    pub struct lpb_ctx {
        _private: [u8; 0],
    }
    #[repr(C)]
    #[derive(Clone, Copy)]
    pub struct lpb_options {
        pub tol: f64,
        pub disp: i32,
        pub ip: i32,
        pub solver_type: i32,
        pub reserved: i32,
        pub alpha0: f64,
        pub max_iter: i64,
    }
    #[repr(C)]
    #[derive(Default, Clone, Copy)]
    pub struct lpb_residual_scalars {
        pub nrm_rp: f64,
        pub nrm_rd: f64,
        pub cx: f64,
        pub by: f64,
        pub xz: f64,
    }
    #[repr(C)]
    #[derive(Default, Clone, Copy)]
    pub struct lpb_direction_in {
        pub corrector: i32,
        pub ip: i32,
        pub eta: f64,
        pub gamma: f64,
        pub mu: f64,
        pub alpha: f64,
    }
    #[repr(C)]
    #[derive(Default, Clone, Copy)]
    pub struct lpb_direction_out {
        pub cu: f64,
        pub bv: f64,
        pub cp: f64,
        pub bq: f64,
        pub nan_pq: i32,
        pub reserved: i32,
    }
    pub const LPB_OK: c_int = 0;
    pub const LPB_ERR_NUMERICAL_PROBLEM: c_int = 2;
    pub const LPB_MEM_HOST: c_int = 0;

    extern "C" {
        pub fn lpb_create(ctx: *mut *mut lpb_ctx, m: i64, n: i64, A: *const f64, lda: i64, b: *const f64,
                          c: *const f64, c0: f64, mem: c_int, stream: *mut c_void) -> c_int;
        pub fn lpb_destroy(ctx: *mut lpb_ctx) -> c_int;
        pub fn lpb_blind_start(ctx: *mut lpb_ctx) -> c_int;
        pub fn lpb_residuals(ctx: *mut lpb_ctx, tau: f64, kappa: f64, out: *mut lpb_residual_scalars) -> c_int;
        pub fn lpb_form_and_factor(ctx: *mut lpb_ctx) -> c_int;
        pub fn lpb_direction(ctx: *mut lpb_ctx, din: *const lpb_direction_in, tau: f64, kappa: f64,
                             out: *mut lpb_direction_out) -> c_int;
        pub fn lpb_assemble_delta(ctx: *mut lpb_ctx, d_tau: f64, alpha_xz: *mut f64) -> c_int;
        pub fn lpb_do_step(ctx: *mut lpb_ctx, alpha: f64, ip: c_int) -> c_int;
        pub fn lpb_extract_x(ctx: *mut lpb_ctx, tau: f64, x_out: *mut f64, fun: *mut f64) -> c_int;
        pub fn lpb_solve(ctx: *mut lpb_ctx, opts: *const lpb_options, x_out: *mut f64, fun: *mut f64,
                         iterations: *mut i64) -> c_int;
        pub fn lpb_last_error() -> *const c_char;
    }
}

// `error.rs`, `float.rs`, `linear_program.rs`, `solvers/mod.rs` are taken from the reference crate
// verbatim by path in a real build (they are host-only and need no change); they are not copied here.
// What follows is the replacement body of `impl Solver<f64> for InteriorPoint<f64>`
// (reference: solvers/interior_point/mod.rs:161-240), shown against the reference's types.
//
// fn solve_normal_form(&self, problem: &Problem<f64>) -> Result<(Array1<f64>, usize), LinearProgramError<f64>> {
//     let (m, n) = problem.A().dim();
//     let a = problem.A().as_standard_layout();                       // row-major, as build() made it
//     let mut ctx = std::ptr::null_mut();
//     check(unsafe { ffi::lpb_create(&mut ctx, m as i64, n as i64, a.as_ptr(), n as i64,
//                                    problem.b().as_ptr(), problem.c().as_ptr(), problem.c0(),
//                                    ffi::LPB_MEM_HOST, std::ptr::null_mut()) })?;
//     let guard = CtxGuard(ctx);                                      // lpb_destroy on drop
//     let (mut tau, mut kappa) = (1.0, 1.0);                          // feasible_point.rs:29-30
//     check(unsafe { ffi::lpb_blind_start(ctx) })?;
//     let mut rs = ffi::lpb_residual_scalars::default();
//     check(unsafe { ffi::lpb_residuals(ctx, tau, kappa, &mut rs) })?;
//     let ini = InitialResiduals::from(&rs, tau, kappa, n);           // residual.rs:13-44
//     let mut ip = self.ip;
//     for iteration in 1..=self.max_iter {                            // mod.rs:213
//         let (mut gamma, mut eta) = if ip { (1.0, 1.0) } else { (0.0, 1.0) };          // feasible_point.rs:119-120
//         let r_g = rs.cx - rs.by + kappa;                            // :124
//         let mu = (rs.xz + tau * kappa) / (n + 1) as f64;            // :125
//         match unsafe { ffi::lpb_form_and_factor(ctx) } {            // newton_equations.rs:48-64
//             ffi::LPB_OK => {}
//             ffi::LPB_ERR_NUMERICAL_PROBLEM => return Err(LinearProgramError::NumericalProblem),
//             e => return Err(device_error(e)),
//         }
//         let mut din = ffi::lpb_direction_in { corrector: 0, ip: ip as i32, eta, gamma, mu, alpha: 0.0 };
//         let mut dout = ffi::lpb_direction_out::default();
//         check(unsafe { ffi::lpb_direction(ctx, &din, tau, kappa, &mut dout) })?;    // rhat.rs:17-35 + sym_solve
//         if dout.nan_pq != 0 { return Err(LinearProgramError::NumericalProblem); }   // newton_equations.rs:190-194
//         let (mut d_tau, mut d_kappa) = delta_scalars(r_g * eta, gamma * mu - tau * kappa, tau, kappa, &dout);
//         let mut axz = [1.0f64; 2];
//         check(unsafe { ffi::lpb_assemble_delta(ctx, d_tau, axz.as_mut_ptr()) })?;    // delta.rs:33-37 + ratio test
//         let alpha = step_size(axz, tau, d_tau, kappa, d_kappa, 1.0);                // feasible_point.rs:134
//         gamma = update_gamma(ip, alpha); eta = if ip { 1.0 } else { 1.0 - gamma };  // :135-136
//         let tk = if ip { (1.0 - alpha) * gamma * mu - tau * kappa - alpha * alpha * d_tau * d_kappa }
//                  else  { gamma * mu - tau * kappa - d_tau * d_kappa };              // rhat.rs:51-66
//         din = ffi::lpb_direction_in { corrector: 1, ip: ip as i32, eta, gamma, mu, alpha };
//         check(unsafe { ffi::lpb_direction(ctx, &din, tau, kappa, &mut dout) })?;
//         (d_tau, d_kappa) = delta_scalars(r_g * eta, tk, tau, kappa, &dout);
//         check(unsafe { ffi::lpb_assemble_delta(ctx, d_tau, axz.as_mut_ptr()) })?;
//         let alpha = if ip { 1.0 } else { step_size(axz, tau, d_tau, kappa, d_kappa, self.alpha0) };  // mod.rs:216-221
//         check(unsafe { ffi::lpb_do_step(ctx, alpha, ip as i32) })?;                 // feasible_point.rs:76-106
//         tau += d_tau * alpha; kappa += d_kappa * alpha;
//         if ip { tau = tau.max(1.0); kappa = kappa.max(1.0); }
//         ip = false;
//         check(unsafe { ffi::lpb_residuals(ctx, tau, kappa, &mut rs) })?;
//         let indicators = Indicators::from_scalars(&rs, &ini, tau, kappa, n, problem.c0());   // indicators.rs:37-55
//         if self.disp { println!("{alpha:3.8}\t{indicators}"); }
//         match indicators.status(tau, kappa, self.tol) {                             // indicators.rs:66-83
//             Status::Optimal => return Ok((extract_x(ctx, tau, n)?, iteration)),    // mod.rs:231
//             Status::Infeasible => return Err(LinearProgramError::Infeasible),
//             Status::Unbounded => return Err(LinearProgramError::Unbounded),
//             Status::Unfinished => {}
//         }
//     }
//     Err(LinearProgramError::IterationLimitExceeded(extract_x(ctx, tau, n)?))       // mod.rs:237-239
// }
