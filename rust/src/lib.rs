//! UNBUILT in this image (no Rust toolchain) -- source-complete Rust host side of the B200 drop-in.
//!
//! Same crate name, module tree and public surface as the reference (`/root/reference/src/lib.rs:71-75`,
//! `prelude.rs:3-11`):
//!
//! ```ignore
//! use ripped::prelude::*;
//! let problem = Problem::target(&c).ub(&A_ub, &b_ub).eq(&A_eq, &b_eq).build()?;
//! let res = InteriorPoint::default().solve(&problem)?;      // res.x(), res.fun(), res.iteration()
//! ```
//!
//! What differs from the reference is only what sits behind `Solver::solve`: the host keeps `tau`, `kappa` and
//! every control-flow decision of `solve_normal_form` (`solvers/interior_point.rs`), and drives the iteration
//! through the phase calls of `include/lpb200.h` (`ffi.rs`); `A`, `M`, its Cholesky factor and the iterate live in
//! HBM.  There is no CPU fallback and no backend feature: the one backend is `liblpb200.so`, built by `build.rs`
//! with nvcc for sm_100a.  `Problem<F>` / `InteriorPoint<F>` exist for `F = f64` and `F = f32` like the reference's (`float.rs`); the arithmetic
//! is FP64 in both cases (an `f32` problem is widened on upload, its result narrowed on the way back).
#![deny(unsafe_code)] // like the reference's lint set (.cargo/config.toml:6); `ffi` and its two callers opt out
#![allow(non_snake_case)]

pub mod error;
pub mod ffi;
pub mod float;
pub mod linear_program;
pub mod prelude;
pub mod solvers;
