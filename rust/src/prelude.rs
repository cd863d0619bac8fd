//! Commonly used imports; use as `use ripped::prelude::*` (mirrors `/root/reference/src/prelude.rs:3-11`).
pub use crate::error::LinearProgramError;
pub use crate::linear_program::Problem;
pub use crate::solvers::interior_point::EquationSolverType;
pub use crate::solvers::interior_point::InteriorPoint;
pub use crate::solvers::Solver;
