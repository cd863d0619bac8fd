"""Iteration count, final x and per-iteration step lengths of a workload for combinations of K1's summation order
("syrk_chain", "syrk_flush_blocks") and the refinement steps ("refine", "refine_max"), against the committed oracle
fixture.     python tools/sweep_refine.py C3 [C2]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SIZES = {"C1": (512, 1024), "C2": (4096, 8192), "C3": (16384, 32768)}


def main():
    import lp_b200
    from lp_b200.api import ResidentProblem
    from bench import synthetic_lp
    for wl in sys.argv[1:]:
        m, n = SIZES[wl]
        c, A_ub, b_ub, A_eq, b_eq = synthetic_lp(m, n, 0)
        pb = lp_b200.Problem.target(c).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
        del A_ub, A_eq
        g = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_%s_seed0.json" % wl)))
        gx = np.load(os.path.join(ROOT, "tests", "golden", g["x_file"]))
        galpha = np.array([t["alpha"] for t in g["trace"]])
        gmu = np.array([t["rho_mu"] for t in g["trace"]])
        print("==== %s oracle: it=%d  alpha[-8:]=%s" % (wl, g["iterations"], np.round(galpha[-8:], 4)), flush=True)
        combos = []
        for chain, fb in ((1, 128), (0, 128), (0, 32)):
            for refine, rmax in ((0, 0), (1, 0), (2, 0), (1, 3)):
                combos.append(dict(syrk_chain=chain, syrk_flush_blocks=fb, refine=refine, refine_max=rmax))
        with ResidentProblem(pb) as rp:
            for combo in combos:
                for k, v in combo.items():
                    rp.set_option(k, v)
                try:
                    res = lp_b200.InteriorPoint.custom().max_iter(60).build().solve_resident(rp)
                    tr = rp.trace().copy()
                    k = min(len(tr), len(galpha))
                    da = np.abs(tr[:k, 0] - galpha[:k])
                    first = int(np.argmax(da > 1e-3)) + 1 if (da > 1e-3).any() else 0
                    print("  %-70s it=%2d max|dx| %.3e fun rel %.1e  first |d alpha| > 1e-3 at it %d (rho_mu %.1e)  alpha[-6:]=%s" % (
                        json.dumps(combo), res.iteration(), np.abs(res.x() - gx).max(),
                        abs(res.fun() - g["fun"]) / abs(g["fun"]), first, gmu[first - 1] if first else 0.0,
                        np.round(tr[-6:, 0], 4)), flush=True)
                except Exception as e:  # noqa: BLE001
                    print("  %-70s %s it=%d" % (json.dumps(combo), type(e).__name__, rp.last_iterations), flush=True)


if __name__ == "__main__":
    main()
