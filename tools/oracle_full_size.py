"""Run the CPU oracle on a full-size seeded workload and write golden fixtures under tests/golden/:

  oracle_<wl>_seed<s>.json          outcome + per-iteration trace at the DEFAULT tolerance (1e-8), plus the
                                    iteration at which each tighter tolerance of TOLS would have stopped
  oracle_<wl>_seed<s>_x.npy         the FULL x (user variables) at the default tolerance
  oracle_<wl>_seed<s>_x_tol1e-10.npy (etc.)  the full x where tol = 1e-9 / 1e-10 stops

`tol` enters the reference only through Indicators::status (indicators.rs:66-83), so ONE run passes through
the iterates at which every tolerance of TOLS stops; they are snapshotted on the way.  (1e-11 is below the
rounding floor of rho_p for these LPs: the reference's own iteration stalls there and M turns singular.)

  python tools/oracle_full_size.py C3 [seed] [--variant splitk]

--variant splitk forms M = A D A^T as two half-K GEMMs added together -- a different, equally valid summation
order: files oracle_<wl>_seed<s>_splitk*.  Comparing it with the default run measures how far two
LAPACK-grade runs of the same algorithm differ on x at each tolerance.  ~40 s per iteration for C3 on
16 cores, ~90 s on 8.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ipm_oracle as o  # noqa: E402

SIZES = {"C1": (512, 1024), "C2": (4096, 8192), "C3": (16384, 32768)}
TOLS = (1e-8, 1e-9, 1e-10)


def splitk_gemm(A, Dinv):
    h = A.shape[1] // 2
    return A[:, :h].dot(Dinv[:h, None] * A[:, :h].T) + A[:, h:].dot(Dinv[h:, None] * A[:, h:].T)


class Done(Exception):
    pass


def main():
    argv = [a for a in sys.argv[1:] if not a.startswith("--")]
    wl = argv[0]
    seed = int(argv[1]) if len(argv) > 1 and "--variant" not in sys.argv[1:3] else 0
    variant = sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else ""
    m, n = SIZES[wl]
    t0 = time.time()
    pb = o.build_problem(*o.synthetic_lp(m, n, seed))
    tr = []
    snaps = {}

    def on_iteration(iteration, pt, ind):
        print("it %d  %.0fs  rho_p %.2e rho_d %.2e rho_A %.2e rho_mu %.2e" % (
            iteration, time.time() - t0, ind.rho_p, ind.rho_d, ind.rho_A, ind.rho_mu), flush=True)
        for tol in TOLS:
            if tol not in snaps and ind.status(pt.tau, pt.kappa, tol) == "Optimal":
                xs = pt.x / pt.tau
                snaps[tol] = dict(x=pb.denormalize_x(xs), fun=pb.denormalize_target(xs), iterations=iteration)
        if len(snaps) == len(TOLS) or (TOLS[0] in snaps and iteration >= snaps[TOLS[0]]["iterations"] + 4):
            raise Done()

    err = None
    try:
        o.InteriorPoint(tol=1e-300, max_iter=80).solve(pb, trace=tr, on_iteration=on_iteration,
                                                       gemm=splitk_gemm if variant == "splitk" else None)
    except Done:
        pass
    except o.LinearProgramError as e:
        err = type(e).__name__
    tag = "oracle_%s_seed%d%s" % (wl, seed, "_" + variant if variant else "")
    gold = os.path.join(ROOT, "tests", "golden")
    out = {"workload": wl, "m": m, "n": n, "seed": seed, "generator": "oracle.ipm_oracle.synthetic_lp (SURVEY 8d)",
           "oracle_backend": "lapack", "variant": variant or "default", "tol": TOLS[0]}
    if TOLS[0] in snaps:
        s = snaps[TOLS[0]]
        x = s["x"]
        out.update(status="Optimal", iterations=s["iterations"], fun=s["fun"], x_head=[float(v) for v in x[:16]],
                   x_sum=float(x.sum()), x_norm2=float(np.linalg.norm(x)), x_file=tag + "_x.npy")
        np.save(os.path.join(gold, tag + "_x.npy"), x)
    else:
        out.update(status=err or "Unfinished", iterations=len(tr))
    tighter = {}
    for tol in TOLS[1:]:
        if tol in snaps:
            s = snaps[tol]
            fn = "%s_x_tol%g.npy" % (tag, tol)
            np.save(os.path.join(gold, fn), s["x"])
            tighter["%g" % tol] = {"iterations": s["iterations"], "fun": s["fun"], "x_file": fn,
                                   "max_abs_dx_vs_default_tol": float(np.abs(s["x"] - snaps[TOLS[0]]["x"]).max())}
    out["tighter"] = tighter
    out["trace"] = tr  # runs past the default stop: rows beyond `iterations` belong to the tighter tolerances
    out["wall_s"] = time.time() - t0
    json.dump(out, open(os.path.join(gold, tag + ".json"), "w"), indent=1)
    print("wrote", tag, out["status"], out["iterations"], out.get("fun"), {k: (v["iterations"], v["max_abs_dx_vs_default_tol"])
                                                                         for k, v in tighter.items()})


if __name__ == "__main__":
    main()
