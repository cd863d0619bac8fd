"""Run the CPU oracle on a full-size seeded workload (e.g. C3) and write the per-iteration trace +
outcome as a golden fixture (tests/golden/oracle_<workload>_seed<seed>.json).  ~50 min for C3 on 8 cores."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ipm_oracle as o  # noqa: E402

SIZES = {"C1": (512, 1024), "C2": (4096, 8192), "C3": (16384, 32768)}


def main():
    wl = sys.argv[1]
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    m, n = SIZES[wl]
    t0 = time.time()
    pb = o.build_problem(*o.synthetic_lp(m, n, seed))
    tr = []
    out = {"workload": wl, "m": m, "n": n, "seed": seed, "generator": "oracle.ipm_oracle.synthetic_lp (SURVEY 8d)",
           "oracle_backend": "lapack"}
    try:
        res = o.InteriorPoint().solve(pb, trace=tr)
        out.update(status="Optimal", iterations=res.iteration, fun=res.fun,
                   x_head=[float(v) for v in res.x[:16]], x_sum=float(res.x.sum()),
                   x_norm2=float(np.linalg.norm(res.x)))
    except o.LinearProgramError as e:
        out.update(status=type(e).__name__, iterations=len(tr))
    out["trace"] = tr
    out["wall_s"] = time.time() - t0
    path = os.path.join(ROOT, "tests", "golden", "oracle_%s_seed%d.json" % (wl, seed))
    json.dump(out, open(path, "w"), indent=1)
    print("wrote", path, out["status"], out["iterations"], out.get("fun"))


if __name__ == "__main__":
    main()
