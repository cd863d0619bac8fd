//! `extern "C"` declarations of `include/lpb200.h` -- the one module that has to relax the reference's
//! `-Dunsafe_code` (`/root/reference/.cargo/config.toml:6`).  Only the entry points the host loop needs.
#![allow(unsafe_code, non_camel_case_types, missing_docs)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct lpb_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
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
#[derive(Default, Clone, Copy, Debug)]
pub struct lpb_residual_scalars {
    pub nrm_rp: f64,
    pub nrm_rd: f64,
    pub cx: f64,
    pub by: f64,
    pub xz: f64,
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct lpb_direction_in {
    pub corrector: i32,
    pub ip: i32,
    pub eta: f64,
    pub gamma: f64,
    pub mu: f64,
    pub alpha: f64,
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
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
pub const LPB_ERR_ITERATION_LIMIT_EXCEEDED: c_int = 7;
pub const LPB_MEM_HOST: c_int = 0;

extern "C" {
    pub fn lpb_options_validate(o: *const lpb_options) -> c_int;
    pub fn lpb_create(
        ctx: *mut *mut lpb_ctx, m: i64, n: i64, A: *const f64, lda: i64, b: *const f64, c: *const f64, c0: f64,
        mem: c_int, stream: *mut c_void,
    ) -> c_int;
    pub fn lpb_destroy(ctx: *mut lpb_ctx) -> c_int;
    pub fn lpb_blind_start(ctx: *mut lpb_ctx) -> c_int;
    pub fn lpb_residuals(ctx: *mut lpb_ctx, tau: f64, kappa: f64, out: *mut lpb_residual_scalars) -> c_int;
    pub fn lpb_form_and_factor(ctx: *mut lpb_ctx) -> c_int;
    pub fn lpb_direction(
        ctx: *mut lpb_ctx, din: *const lpb_direction_in, tau: f64, kappa: f64, out: *mut lpb_direction_out,
    ) -> c_int;
    pub fn lpb_assemble_delta(ctx: *mut lpb_ctx, d_tau: f64, alpha_xz: *mut f64) -> c_int;
    pub fn lpb_do_step(ctx: *mut lpb_ctx, alpha: f64, ip: c_int) -> c_int;
    pub fn lpb_extract_x(ctx: *mut lpb_ctx, tau: f64, x_out: *mut f64, fun: *mut f64) -> c_int;
    pub fn lpb_last_error() -> *const c_char;
}

/// Detail text of the last failing call on this thread (empty if none).
pub fn last_error() -> String {
    // SAFETY: lpb_last_error returns a pointer to a NUL-terminated thread-local buffer that outlives the call.
    unsafe {
        let p = lpb_last_error();
        if p.is_null() {
            String::new()
        } else {
            CStr::from_ptr(p).to_string_lossy().into_owned()
        }
    }
}

/// Owns one `lpb_ctx` (device memory of one resident problem); `lpb_destroy` on drop, also on every early
/// return of the host loop.  Not `Send`/`Sync`: a context belongs to the thread that drives it; concurrent
/// `solve(&self, &problem)` calls each create their own (the reference's `solve` is re-entrant the same way).
pub struct CtxGuard(*mut lpb_ctx);

impl CtxGuard {
    /// Upload a slack-form problem (row-major `a`, `m x n`, leading dimension `n`).
    pub fn create(m: usize, n: usize, a: &[f64], b: &[f64], c: &[f64], c0: f64) -> Result<Self, c_int> {
        assert!(a.len() == m * n && b.len() == m && c.len() == n);
        let mut ctx: *mut lpb_ctx = std::ptr::null_mut();
        // SAFETY: the slices outlive the call (liblpb200 copies them to the device before returning) and their
        // lengths match the dimensions passed.
        let rc = unsafe {
            lpb_create(&mut ctx, m as i64, n as i64, a.as_ptr(), n as i64, b.as_ptr(), c.as_ptr(), c0, LPB_MEM_HOST,
                       std::ptr::null_mut())
        };
        if rc == LPB_OK { Ok(CtxGuard(ctx)) } else { Err(rc) }
    }

    pub fn raw(&self) -> *mut lpb_ctx {
        self.0
    }
}

impl Drop for CtxGuard {
    fn drop(&mut self) {
        if !self.0.is_null() {
            // SAFETY: the pointer came from lpb_create and is destroyed exactly once.
            unsafe {
                lpb_destroy(self.0);
            }
            self.0 = std::ptr::null_mut();
        }
    }
}
