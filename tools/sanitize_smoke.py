"""Small end-to-end invocations of every single-GPU kernel family, meant to run under
`compute-sanitizer --tool memcheck` (one tool per gpurun call, see B200_PROFILING.md)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lp_b200  # noqa: E402
from oracle import ipm_oracle as o  # noqa: E402


def build(c, A_ub, b_ub, A_eq, b_eq):
    return lp_b200.Problem.target(c).ub(A_ub, b_ub).eq(A_eq, b_eq).build()


for (m, n, seed) in [(64, 128, 0), (130, 301, 2), (300, 1500, 5)]:
    args = o.synthetic_lp(m, n, seed)
    ref = o.InteriorPoint().solve(o.build_problem(*args))
    res = lp_b200.InteriorPoint.default().solve(build(*args))
    assert abs(res.iteration() - ref.iteration) <= 1 and np.abs(res.x() - ref.x).max() < 1e-6
    print("solve", m, n, "ok", res.iteration(), flush=True)
As, bs, cs = [], [], []
for i in range(5):
    pb = o.build_problem(*o.synthetic_lp(10, 30, 1000 + i))
    As.append(pb.A), bs.append(pb.b), cs.append(pb.c)
r = lp_b200.solve_batched(np.stack(As), np.stack(bs), np.stack(cs), n_slack=5)
assert (r.status == 0).all()
As, bs, cs = [], [], []
for i in range(3):
    pb = o.build_problem(*o.synthetic_lp(64, 128, 2000 + i))
    As.append(pb.A), bs.append(pb.b), cs.append(pb.c)
r = lp_b200.solve_batched(np.stack(As), np.stack(bs), np.stack(cs), n_slack=32)
assert (r.status == 0).all()
print("batched ok", flush=True)
