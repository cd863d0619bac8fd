//! Error type of the solvers.  Same seven variants, in the same order and with the same user-facing messages as
//! `/root/reference/src/error.rs:7-29`; the C-ABI status codes 1..7 of `include/lpb200.h` are these variants in
//! declaration order.  `Display` / `Error` are written out by hand (no derive macro dependency).
use ndarray::Array1;
use std::fmt::{self, Debug, Display};

/// Problems encountered while building or solving a linear program.
#[derive(Debug)]
pub enum LinearProgramError<F: Debug> {
    /// No constraint rows at all (status 1).
    Unconstrained,
    /// The factorisation of the normal matrix failed, or p / q contain NaN (status 2).
    NumericalProblem,
    /// A builder parameter is out of range -- or the request cannot run on the B200 path, see the text (status 3).
    InvalidParameter(&'static str),
    /// Shapes of c, A_ub, b_ub, A_eq, b_eq do not agree (status 4).
    IncompatibleInputDimensions,
    /// The homogeneous model certified primal infeasibility (status 5).
    Infeasible,
    /// The homogeneous model certified unboundedness (status 6).
    Unbounded,
    /// `max_iter` iterations without meeting the tolerances; carries the best `x / tau` in SLACK form (status 7).
    IterationLimitExceeded(Array1<F>),
}

impl<F: Debug> Display for LinearProgramError<F> {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        use LinearProgramError::*;
        match self {
            Unconstrained => f.write_str(
                "The problem is unconstrained, meaning the solution is the all-zeros vector if `c` is nonnegative, \
                 or unbounded otherwise.",
            ),
            NumericalProblem => f.write_str(
                "The solver encountered numerical problems it could not recover from. Likely causes are linearly \
                 dependent constraints or variables whose scale differs by multiple orders of magnitude.",
            ),
            InvalidParameter(what) => write!(f, "A parameter was set to an invalid value: {}", what),
            IncompatibleInputDimensions => {
                f.write_str("The dimensions of your cost- and constraint arrays do not align.")
            }
            Infeasible => {
                f.write_str("The solver finished successfully, it appears that the problem is infeasible.")
            }
            Unbounded => {
                f.write_str("The solver finished successfully, it appears that your problem is unbounded.")
            }
            IterationLimitExceeded(x) => write!(
                f,
                "The solver failed to converge within the maximum number of iterations. Best solution after the \
                 final iteration:\n{:#?}",
                x
            ),
        }
    }
}

impl<F: Debug> std::error::Error for LinearProgramError<F> {}

impl<F: Debug> LinearProgramError<F> {
    /// Map a non-zero lpb status code that carries no payload (`include/lpb200.h`): 1..6 are the variants above;
    /// negative codes are device-side failures the reference has no variant for -- they surface as
    /// `InvalidParameter` with a static description (the detail text of `lpb_last_error()` goes to stderr).
    pub(crate) fn from_code(code: i32) -> Self {
        match code {
            1 => Self::Unconstrained,
            2 => Self::NumericalProblem,
            3 => Self::InvalidParameter("rejected by liblpb200 (lpb_options_validate)"),
            4 => Self::IncompatibleInputDimensions,
            5 => Self::Infeasible,
            6 => Self::Unbounded,
            // 7 (IterationLimitExceeded) needs x / tau: built by the caller, never through this function
            7 => Self::NumericalProblem,
            -1 => Self::InvalidParameter("CUDA error on the B200 path (details on stderr)"),
            -2 => Self::InvalidParameter("NCCL error on the B200 path (details on stderr)"),
            -3 => Self::InvalidParameter("no CUDA device: the B200 path has no CPU fallback"),
            -4 => Self::InvalidParameter("bad argument passed to liblpb200"),
            -5 => Self::InvalidParameter(
                "only EquationSolverType::Cholesky runs on the B200 path (Inverse / LeastSquares are CPU fallbacks)",
            ),
            _ => Self::InvalidParameter("unknown liblpb200 status"),
        }
    }
}
