//! What a successful solve returns (reference: `OptimizeResult`, `/root/reference/src/solvers/mod.rs:19-49`).
//! The three accessors keep the reference's names and signatures; everything else is this crate's own.
use ndarray::Array1;
use std::fmt;

/// Solution vector in the user's variables, objective value and iteration count of one solve.
pub struct OptimizeResult<F> {
    solution: Array1<F>,
    objective: F,
    iterations: usize,
}

impl<F> OptimizeResult<F> {
    pub(crate) fn new(solution: Array1<F>, objective: F, iterations: usize) -> Self {
        OptimizeResult { solution, objective, iterations }
    }

    /// The solution (slack variables removed).
    pub fn x(&self) -> &Array1<F> { &self.solution }

    /// Objective value `c'x + c0`.
    pub fn fun(&self) -> &F { &self.objective }

    /// Number of interior-point iterations taken.
    pub fn iteration(&self) -> usize { self.iterations }

    /// Take the result apart: `(x, fun, iteration)`.
    pub fn into_parts(self) -> (Array1<F>, F, usize) { (self.solution, self.objective, self.iterations) }
}

impl<F: fmt::Debug> fmt::Debug for OptimizeResult<F> {
    fn fmt(&self, f: &mut fmt::Formatter<'_>) -> fmt::Result {
        f.debug_struct("OptimizeResult")
            .field("iteration", &self.iterations)
            .field("fun", &self.objective)
            .field("len(x)", &self.solution.len())
            .finish()
    }
}
