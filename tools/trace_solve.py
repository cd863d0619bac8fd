"""Solve a seeded SURVEY-8(d) LP on the GPU and print the per-iteration trace next to the oracle's golden
trace (tests/golden/oracle_<WL>_seed0.json) when there is one.
python tools/trace_solve.py C3 [key=value ...]   (context options, e.g. solve_impl=1)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SIZES = {"C1": (512, 1024), "C2": (4096, 8192), "C3": (16384, 32768)}


def main():
    import lp_b200
    from lp_b200.api import ResidentProblem
    from bench import synthetic_lp
    wl = sys.argv[1]
    m, n = SIZES[wl]
    opts = dict(kv.split("=") for kv in sys.argv[2:])
    c, A_ub, b_ub, A_eq, b_eq = synthetic_lp(m, n, 0)
    pb = lp_b200.Problem.target(c).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
    gold = None
    gp = os.path.join(ROOT, "tests", "golden", "oracle_%s_seed0.json" % wl)
    if os.path.exists(gp):
        gold = json.load(open(gp))
    with ResidentProblem(pb) as rp:
        for k, v in opts.items():
            rp.set_option(k, int(v))
        try:
            res = lp_b200.InteriorPoint.custom().max_iter(60).build().solve_resident(rp)
            print("GPU %s: Optimal it=%d fun=%.12f" % (opts, res.iteration(), res.fun()))
        except Exception as e:  # noqa: BLE001
            print("GPU %s: %s it=%d" % (opts, type(e).__name__, rp.last_iterations))
        tr = rp.trace()
    if gold:
        print("oracle: %s it=%d fun=%.12f" % (gold["status"], gold["iterations"], gold.get("fun", float("nan"))))
    for i, row in enumerate(tr):
        line = "%2d gpu a=%.6f rp=%.4e rd=%.4e rA=%.4e rmu=%.4e tau=%.6f" % (i + 1, row[0], row[1], row[2], row[3], row[5], row[8])
        if gold and i < len(gold["trace"]):
            g = gold["trace"][i]
            line += " | ora a=%.6f rp=%.4e rd=%.4e rA=%.4e rmu=%.4e tau=%.6f" % (
                g["alpha"], g["rho_p"], g["rho_d"], g["rho_A"], g["rho_mu"], g["tau"])
        print(line)
        line = "   gpu kappa=%.6e cp=%.15e bq=%.15e bq-cp=%.6e cu=%.10e bv=%.10e d_tau=%.6e d_kappa=%.6e" % (
            row[9], row[10], row[11], row[11] - row[10], row[12], row[13], row[14], row[15])
        print(line)
        if gold and i < len(gold["trace"]) and "cp" in gold["trace"][i]:
            g = gold["trace"][i]
            print("   ora kappa=%.6e cp=%.15e bq=%.15e bq-cp=%.6e cu=%.10e bv=%.10e d_tau=%.6e d_kappa=%.6e" % (
                g["kappa"], g["cp"], g["bq"], g["bq"] - g["cp"], g["cu"], g["bv"], g["d_tau"], g["d_kappa"]))


if __name__ == "__main__":
    main()
