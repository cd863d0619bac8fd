"""The reference's own known-answer problems.

Each entry cites the reference test it comes from (paths relative to /root/reference).
"""
import numpy as np

GOLDEN = {
    # src/lib.rs:84-113, src/lib.rs:23-51, src/solvers/interior_point/mod.rs:256-273
    "G1": dict(c=[-1.0, 4.0], A_ub=[[-3.0, 1.0], [1.0, 2.0]], b_ub=[6.0, 4.0],
               A_eq=[[1.0, 1.0]], b_eq=[1.0], x=[1.0, 0.0], eps=1e-6),
    # src/solvers/interior_point/mod.rs:181-192
    "G2": dict(c=[-1.0, 4.0], A_ub=[[-3.0, 1.0], [1.0, 2.0]], b_ub=[6.0, 4.0],
               x=[4.0, 0.0], eps=1e-6),
    # src/solvers/interior_point/mod.rs:319-331
    "G3": dict(c=[-1.0, 4.0, -1.2], A_eq=[[2.0, 1.0, 0.0], [0.0, 2.0, 1.0], [1.0, 0.0, 2.0]],
               b_eq=[1.0, 2.0, 3.0], x=[1.0 / 3.0, 1.0 / 3.0, 4.0 / 3.0], eps=1e-6),
    # src/solvers/interior_point/mod.rs:332-344
    "G4": dict(c=[-1.0, 4.0, -1.2], A_ub=[[2.0, 1.0, 0.0], [0.0, 2.0, 1.0], [1.0, 0.0, 2.0]],
               b_ub=[1.0, 2.0, 3.0], x=[0.5, 0.0, 1.25], eps=1e-6),
}


def golden_arrays(name):
    g = GOLDEN[name]
    def arr(k):
        return None if k not in g else np.asarray(g[k], dtype=np.float64)
    return arr("c"), arr("A_ub"), arr("b_ub"), arr("A_eq"), arr("b_eq"), np.asarray(g["x"]), g["eps"]


def symmetric_example(N=1000):
    """examples/symmetric.rs:10-25: A_ub = 1 - I, b_ub = N-1, c = -1; x == 1 to 1e-10."""
    A_ub = np.ones((N, N)) - np.eye(N)
    b_ub = np.full(N, float(N - 1))
    c = -np.ones(N)
    return c, A_ub, b_ub, None, None, np.ones(N), 1e-10
