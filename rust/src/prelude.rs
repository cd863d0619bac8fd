//! `use ripped::prelude::*` brings the same five names into scope as the reference's prelude
//! (`/root/reference/src/prelude.rs:3-11`).
pub use crate::{
    error::LinearProgramError,
    linear_program::Problem,
    solvers::{
        interior_point::{EquationSolverType, InteriorPoint},
        Solver,
    },
};
