//! Solvers for linear programs: the `Solver` trait and `OptimizeResult` of `/root/reference/src/solvers/mod.rs:12-49`.
pub mod interior_point;

pub use interior_point::InteriorPoint;
use ndarray::Array1;
use std::fmt::Debug;

use crate::{error::LinearProgramError, linear_program::Problem};

/// The plug-in point of the crate: "any solver should implement this" (reference `solvers/mod.rs:11-16`).
pub trait Solver<F: Debug> {
    /// Solve a linear program; errors are the `LinearProgramError` variants, never a panic.
    fn solve(&self, problem: &Problem<F>) -> Result<OptimizeResult<F>, LinearProgramError<F>>;
}

/// Outcome of a successful solve.
pub struct OptimizeResult<F> {
    x: Array1<F>,
    fun: F,
    iteration: usize,
}

impl<F> OptimizeResult<F> {
    pub(crate) fn new(x: Array1<F>, fun: F, iteration: usize) -> Self {
        Self { x, fun, iteration }
    }

    /// Number of interior-point iterations taken.
    pub fn iteration(&self) -> usize {
        self.iteration
    }

    /// Objective value `c'x + c0`.
    pub fn fun(&self) -> &F {
        &self.fun
    }

    /// The solution in the user's variables (slack variables removed).
    pub fn x(&self) -> &Array1<F> {
        &self.x
    }
}
