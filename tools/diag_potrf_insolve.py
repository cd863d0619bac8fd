"""In-solve determinism check of K2: solve a seeded LP with option potrf_verify (every M factored twice).
python tools/diag_potrf_insolve.py M N [trsm_impl update_impl verify_mode]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import lp_b200
    from lp_b200.api import ResidentProblem
    from bench import synthetic_lp
    m, n = int(sys.argv[1]), int(sys.argv[2])
    ti, ui, mode = (int(v) for v in (sys.argv[3:6] if len(sys.argv) > 5 else (0, 0, 1)))
    c, A_ub, b_ub, A_eq, b_eq = synthetic_lp(m, n, 0)
    pb = lp_b200.Problem.target(c).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
    solver = lp_b200.InteriorPoint.custom().max_iter(40).build()
    with ResidentProblem(pb) as rp:
        rp.set_option("trsm_impl", ti)
        rp.set_option("update_impl", ui)
        rp.set_option("potrf_verify", mode)
        for attempt in range(2):
            try:
                res = solver.solve_resident(rp)
                print("variant trsm=%d update=%d mode=%d attempt %d: Optimal it=%d fun=%.10f" % (
                    ti, ui, mode, attempt, res.iteration(), res.fun()), flush=True)
            except Exception as e:  # noqa: BLE001
                print("variant trsm=%d update=%d mode=%d attempt %d: %s it=%d" % (
                    ti, ui, mode, attempt, type(e).__name__, rp.last_iterations), flush=True)


if __name__ == "__main__":
    main()
