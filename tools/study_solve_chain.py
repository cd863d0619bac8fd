"""Offline (CPU) stability study for a faster triangular-solve chain (DESIGN.md section 9, row "solves").

K3's critical path per 128-row block j is   w_j = inv(L_jj) (c_j' - L[j][j-1] w_{j-1})   (a block product, then the
blocked substitution with L_jj: ~5 us).  With pre-multiplied sub-diagonal blocks  G_j = inv(L_jj) L[j][j-1]  the
chain step becomes  w_j = u_j - G_j w_{j-1}  with  u_j = inv(L_jj) c_j'  computed OFF the chain: one 128 x 128 GEMV.
Mathematically identical, numerically not: this script runs the oracle's interior-point iteration with that solve
(and the mirrored backward sweep) and reports iteration counts / objective / x against the plain LAPACK solve,
with and without the one step of iterative refinement the GPU path applies (lpb_api.cu: direction()).

    python tools/study_solve_chain.py [m n seed]...
"""
import os
import sys

import numpy as np
from scipy.linalg import solve_triangular

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ipm_oracle as o  # noqa: E402

NB = 128


class ChainSolver(o.EquationsSolver):
    refine = 0

    def __init__(self, pt, pb, backend="lapack", gemm=None):
        super().__init__(pt, pb, "lapack", gemm)
        self.L = np.tril(self.factor.T)
        m = self.L.shape[0]
        self.blocks = [(s, min(s + NB, m)) for s in range(0, m, NB)]
        L = self.L
        self.G, self.H = {}, {}
        for j, (a, b) in enumerate(self.blocks):
            if j > 0:
                pa, pb_ = self.blocks[j - 1]
                self.G[j] = solve_triangular(L[a:b, a:b], L[a:b, pa:pb_], lower=True)           # inv(L_jj) L[j][j-1]
            if j + 1 < len(self.blocks):
                na, nb_ = self.blocks[j + 1]
                self.H[j] = solve_triangular(L[a:b, a:b].T, L[na:nb_, a:b].T, lower=False)       # inv(L_jj)^T L[j+1][j]^T

    def _chain(self, r):
        L, blocks = self.L, self.blocks
        w = np.zeros_like(r)
        for j, (a, b) in enumerate(blocks):
            cj = r[a:b].copy()
            if j > 1:
                pa = blocks[j - 1][0]
                cj -= L[a:b, :pa].dot(w[:pa])                        # everything but the previous block: off the chain
            u = solve_triangular(L[a:b, a:b], cj, lower=True)
            w[a:b] = u - (self.G[j].dot(w[blocks[j - 1][0]:blocks[j - 1][1]]) if j > 0 else 0.0)
        x = np.zeros_like(r)
        for j in range(len(blocks) - 1, -1, -1):
            a, b = blocks[j]
            cj = w[a:b].copy()
            if j + 2 < len(blocks):
                nb2 = blocks[j + 2][0]
                cj -= L[nb2:, a:b].T.dot(x[nb2:])
            u = solve_triangular(L[a:b, a:b].T, cj, lower=False)
            x[a:b] = u - (self.H[j].dot(x[blocks[j + 1][0]:blocks[j + 1][1]]) if j + 1 < len(blocks) else 0.0)
        return x

    def solve(self, r):
        return self._chain(r)

    def sym_solve(self, A, r1, r2):
        u, v = super().sym_solve(A, r1, r2)
        for _ in range(self.refine):                                  # residual against the operator A Dinv A^T
            rho = r2 - A.dot(u)
            v = v + self.solve(rho)
            u = self.Dinv * (A.T.dot(v) - r1)
        return u, v


class PlainRefined(o.EquationsSolver):
    refine = 0

    def sym_solve(self, A, r1, r2):
        u, v = super().sym_solve(A, r1, r2)
        for _ in range(self.refine):
            rho = r2 - A.dot(u)
            v = v + self.solve(rho)
            u = self.Dinv * (A.T.dot(v) - r1)
        return u, v


def run(cls, refine, pb):
    saved = o.EquationsSolver
    cls.refine = refine
    o.EquationsSolver = cls
    try:
        return o.InteriorPoint().solve(pb)
    finally:
        o.EquationsSolver = saved


def main(cases, golden_only=False):
    import json
    for (m, n, seed) in cases:
        pb = o.build_problem(*o.synthetic_lp(m, n, seed))
        variants = [("plain + 1 refinement", PlainRefined, 1), ("G/H chain", ChainSolver, 0),
                    ("G/H chain + 1 refinement", ChainSolver, 1)]
        if golden_only:  # full-size configs: the plain run is the committed fixture (tools/oracle_full_size.py)
            name = {(512, 1024): "C1", (4096, 8192): "C2", (16384, 32768): "C3"}[(m, n)]
            g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden",
                                            "oracle_%s_seed%d.json" % (name, seed))))
            ref_it, ref_fun, ref_x = g["iterations"], g["fun"], None
            variants = variants[2:]
        else:
            ref = o.InteriorPoint().solve(pb)
            ref_it, ref_fun, ref_x = ref.iteration, ref.fun, ref.x
        print("m=%d n=%d seed=%d  plain LAPACK: %d iterations, fun=%.12g" % (m, n, seed, ref_it, ref_fun), flush=True)
        for name, cls, refine in variants:
            try:
                r = run(cls, refine, pb)
                dx = np.abs(r.x - ref_x).max() if ref_x is not None else float("nan")
                print("   %-26s %d iterations, fun rel diff %.2e, max |dx| %.2e" % (
                    name, r.iteration, abs(r.fun - ref_fun) / abs(ref_fun), dx), flush=True)
            except o.LinearProgramError as e:
                print("   %-26s FAILED: %s" % (name, type(e).__name__), flush=True)


if __name__ == "__main__":
    golden = "--golden" in sys.argv
    a = [int(v) for v in sys.argv[1:] if v != "--golden"]
    cases = [tuple(a[i:i + 3]) for i in range(0, len(a), 3)] or [(512, 1024, 0), (1024, 2048, 1), (2048, 4096, 0)]
    main(cases, golden)
