"""Where the end-to-end overhead of InteriorPoint.solve(problem) goes: context creation (allocations, H2D of A, the
structure scan), the solve, teardown.  Run with LPB_TIME_CREATE=1 for the library's own stage lines.
    python tools/time_create.py C3"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SIZES = {"C1": (512, 1024), "C2": (4096, 8192), "C3": (16384, 32768)}


def main():
    import torch
    import lp_b200
    from lp_b200.api import ResidentProblem
    from bench import synthetic_lp
    m, n = SIZES[sys.argv[1]]
    c, A_ub, b_ub, A_eq, b_eq = synthetic_lp(m, n, 0)
    pb = lp_b200.Problem.target(c).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
    print("pinned host buffers:", type(pb.A()).__name__)
    solver = lp_b200.InteriorPoint.default()
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rp = ResidentProblem(pb)
        t1 = time.perf_counter()
        res = solver.solve_resident(rp)
        t2 = time.perf_counter()
        rp.close()
        t3 = time.perf_counter()
        print("rep %d: create %.1f ms, solve %.1f ms (%d it), close %.1f ms, total %.1f ms" % (
            rep, (t1 - t0) * 1e3, (t2 - t1) * 1e3, res.iteration(), (t3 - t2) * 1e3, (t3 - t0) * 1e3), flush=True)
    t0 = time.perf_counter()
    res = solver.solve(pb)
    print("InteriorPoint.solve(problem): %.1f ms" % ((time.perf_counter() - t0) * 1e3))


if __name__ == "__main__":
    main()
