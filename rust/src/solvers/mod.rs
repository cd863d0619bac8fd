//! Solvers for linear programs: the plug-in trait of the crate (reference `/root/reference/src/solvers/mod.rs:11-16`:
//! "any solver should implement this") and its one implementation here, the B200 interior-point solver.
pub mod interior_point;
mod result;

pub use interior_point::InteriorPoint;
pub use result::OptimizeResult;

use crate::{error::LinearProgramError, linear_program::Problem};
use std::fmt::Debug;

/// A solver maps a slack-form [`Problem`] to an [`OptimizeResult`] or to one of the [`LinearProgramError`]s;
/// it never panics on the solve path.
pub trait Solver<F: Debug> {
    /// Solve `problem`.
    fn solve(&self, problem: &Problem<F>) -> Result<OptimizeResult<F>, LinearProgramError<F>>;
}
