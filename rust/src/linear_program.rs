//! Definition of a linear program in slack form -- same public surface as
//! `/root/reference/src/linear_program.rs` (`Problem::target`, `ProblemBuilder::{new, ub, eq, build}`,
//! accessors `A() b() c()`); host-only, O(mn), runs once per problem.
//!
//! ```text
//! min_x c ' x   st   A x == b,  x >= 0          A = [[A_ub, I], [A_eq, 0]],  b = [b_ub; b_eq],  c = [c; 0]
//! ```
use crate::error::LinearProgramError;
use crate::float::Float;
use ndarray::{s, Array1, Array2};

/// A linear program in slack form (only equality constraints, `x >= 0`).
pub struct Problem<F> {
    a_slack: Array2<F>,   // [[A_ub, I], [A_eq, 0]], row-major
    b_slack: Array1<F>,   // [b_ub; b_eq]
    c_slack: Array1<F>,   // [c; 0]
    offset: F,            // constant term of the objective (0 for every problem the builder makes)
    slack_count: usize,   // number of inequality rows = number of trailing slack variables
}

impl<F: Float> Problem<F> {
    /// Start the builder with the cost vector `c` of `min c'x` (reference `linear_program.rs:37-39`).
    pub fn target(c: &Array1<F>) -> ProblemBuilder<'_, F> {
        ProblemBuilder::new(c)
    }
}

impl<F: Float> Problem<F> {
    /// The slack-form constraint matrix (row-major, `m x n`).
    pub fn A(&self) -> &Array2<F> { &self.a_slack }

    /// The slack-form right-hand side.
    pub fn b(&self) -> &Array1<F> { &self.b_slack }

    /// The slack-form cost vector (`c` followed by `n_slack` zeros).
    pub fn c(&self) -> &Array1<F> { &self.c_slack }

    pub(crate) fn c0(&self) -> F { self.offset }

    pub(crate) fn n_slack(&self) -> usize { self.slack_count }

    /// Drop the slack variables again (reference `linear_program.rs:65-69`).
    pub(crate) fn denormalize_x_into(&self, x_slack: Array1<F>) -> Array1<F> {
        let keep = x_slack.len() - self.slack_count;
        x_slack.slice(s![..keep]).to_owned()
    }
}

/// Collects borrowed constraint blocks and converts them to slack form in `build`.
pub struct ProblemBuilder<'a, F> {
    cost: &'a Array1<F>,
    upper: Option<(&'a Array2<F>, &'a Array1<F>)>,   // A_ub x <= b_ub
    equal: Option<(&'a Array2<F>, &'a Array1<F>)>,   // A_eq x == b_eq
}

impl<'a, F: Float> ProblemBuilder<'a, F> {
    /// Start building a problem with cost vector `c`.
    pub fn new(c: &'a Array1<F>) -> Self {
        ProblemBuilder { cost: c, upper: None, equal: None }
    }

    /// Inequality block `A x <= b`.
    pub fn ub(mut self, A: &'a Array2<F>, b: &'a Array1<F>) -> Self {
        self.upper = Some((A, b));
        self
    }

    /// Equality block `A x == b`.
    pub fn eq(mut self, A: &'a Array2<F>, b: &'a Array1<F>) -> Self {
        self.equal = Some((A, b));
        self
    }

    /// Validate the shapes and assemble the slack form (reference `linear_program.rs:125-169`):
    /// `Unconstrained` without any row, `IncompatibleInputDimensions` on a shape mismatch.
    pub fn build(self) -> Result<Problem<F>, LinearProgramError<F>> {
        let n_c = self.cost.len();
        let (rows_ub, cols_ub, len_b_ub) = match self.upper {
            Some((A, b)) => (A.nrows(), A.ncols(), b.len()),
            None => (0, n_c, 0), // (0, n) placeholders
        };
        let (rows_eq, cols_eq, len_b_eq) = match self.equal {
            Some((A, b)) => (A.nrows(), A.ncols(), b.len()),
            None => (0, n_c, 0),
        };
        if rows_ub + rows_eq == 0 {
            return Err(LinearProgramError::Unconstrained);
        }
        if cols_ub != cols_eq || cols_eq != n_c || rows_ub != len_b_ub || rows_eq != len_b_eq {
            return Err(LinearProgramError::IncompatibleInputDimensions);
        }
        let (m, n) = (rows_ub + rows_eq, n_c + rows_ub);
        let (zero, one) = (F::from_f64(0.0), F::from_f64(1.0));
        let mut A = Array2::<F>::from_elem((m, n), zero);
        let mut b = Array1::<F>::from_elem(m, zero);
        let mut c = Array1::<F>::from_elem(n, zero);
        if let Some((A_ub, b_ub)) = self.upper {
            A.slice_mut(s![..rows_ub, ..n_c]).assign(A_ub);
            for i in 0..rows_ub {
                A[[i, n_c + i]] = one; // the slack block [I; 0]
            }
            b.slice_mut(s![..rows_ub]).assign(b_ub);
        }
        if let Some((A_eq, b_eq)) = self.equal {
            A.slice_mut(s![rows_ub.., ..n_c]).assign(A_eq);
            b.slice_mut(s![rows_ub..]).assign(b_eq);
        }
        c.slice_mut(s![..n_c]).assign(self.cost);
        Ok(Problem { a_slack: A, b_slack: b, c_slack: c, offset: zero, slack_count: rows_ub })
    }
}
