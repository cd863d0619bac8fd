"""What stops a run at a tolerance tighter than the default?  Per option combination: outcome at tol = 1e-10 and the
last rows of the trace (alpha, rho_p, rho_d, rho_g, rho_mu) next to the oracle's.   python tools/sweep_tight.py C2 [C1]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SIZES = {"C1": (512, 1024), "C2": (4096, 8192), "C3": (16384, 32768)}
COLS = ["alpha", "rho_p", "rho_d", "rho_A", "rho_g", "rho_mu"]


def main():
    import lp_b200
    from lp_b200.api import ResidentProblem
    from bench import synthetic_lp
    tol = float(os.environ.get("TIGHT_TOL", "1e-10"))
    for wl in sys.argv[1:]:
        m, n = SIZES[wl]
        c, A_ub, b_ub, A_eq, b_eq = synthetic_lp(m, n, 0)
        pb = lp_b200.Problem.target(c).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_%s_seed0.json" % wl)))
        print("==== %s tol %.0e; oracle trace tail:" % (wl, tol))
        for k, t in enumerate(g["trace"]):
            if k >= len(g["trace"]) - 5:
                print("   oracle it %2d " % (k + 1) + " ".join("%s %.3e" % (c_, t[c_]) for c_ in COLS))
        base = dict(refine=0, refine_max=0, syrk_chain=0, syrk_flush_blocks=32, solve_impl=0, trsm_impl=0, update_impl=0,
                    potf2_impl=0, syrk_impl=0, potrf_lookahead=1)
        combos = [dict(), dict(refine=1), dict(refine=2), dict(solve_impl=2), dict(solve_impl=3), dict(trsm_impl=1),
                  dict(potf2_impl=1, solve_impl=2), dict(potf2_impl=1, solve_impl=2, update_impl=1),
                  dict(potf2_impl=1, solve_impl=2, update_impl=1, syrk_impl=1), dict(syrk_chain=1),
                  dict(syrk_flush_blocks=128), dict(potrf_lookahead=0)]
        with ResidentProblem(pb) as rp:
            for combo in combos:
                for k, v in dict(base, **combo).items():
                    rp.set_option(k, v)
                try:
                    res = lp_b200.InteriorPoint.custom().tol(tol).max_iter(40).build().solve_resident(rp)
                    out = "Optimal it=%d" % res.iteration()
                except Exception as e:  # noqa: BLE001
                    out = "%s it=%d" % (type(e).__name__, rp.last_iterations)
                tr = rp.trace()
                print("  %-60s %s" % (json.dumps(combo), out))
                for k in range(max(0, len(tr) - 7), len(tr)):
                    print("      it %2d " % (k + 1) + " ".join("%s %.3e" % (c_, tr[k][i]) for i, c_ in enumerate(COLS)), flush=True)


if __name__ == "__main__":
    main()
