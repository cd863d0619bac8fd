//! The float types a problem can be stated in: `f64` and `f32`, as in the reference (`/root/reference/src/float.rs:43`).
//! On the B200 path every problem is SOLVED in FP64 (the kernels are FP64 tensor-core code; the parity bar of the
//! path is FP64): an `f32` problem is widened when it is uploaded and its result narrowed when it comes back.
use std::fmt::Debug;

/// `f32` or `f64`.
pub trait Float: Copy + Debug + PartialOrd + 'static {
    /// Widen to the arithmetic type of the device path.
    fn to_f64(self) -> f64;
    /// Narrow a device result to the caller's type.
    fn from_f64(v: f64) -> Self;
}

impl Float for f64 {
    fn to_f64(self) -> f64 {
        self
    }
    fn from_f64(v: f64) -> Self {
        v
    }
}

impl Float for f32 {
    fn to_f64(self) -> f64 {
        self as f64
    }
    fn from_f64(v: f64) -> Self {
        v as f32
    }
}
