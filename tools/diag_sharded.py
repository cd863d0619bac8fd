"""Diagnose rank divergence of the column-sharded path: run under torchrun, compare per-iteration
traces and buffer checksums across ranks.   torchrun ... tools/diag_sharded.py M N [device]"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import lp_b200
    from lp_b200.api import ShardedProblem, SyntheticShardedProblem
    m, n = int(sys.argv[1]), int(sys.argv[2])
    device_gen = len(sys.argv) > 3 and sys.argv[3] == "device"
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group(backend="nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    if device_gen:
        rp = SyntheticShardedProblem(m, n, 0, rank, world, dist)
    else:
        sys.path.insert(0, ROOT)
        from bench import synthetic_lp
        c, A_ub, b_ub, A_eq, b_eq = synthetic_lp(m, n, 0)
        pb = lp_b200.Problem.target(c).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
        rp = ShardedProblem(pb, rank, world, dist)
    solver = lp_b200.InteriorPoint.custom().max_iter(60).build()
    rp.set_option("check_replicas", 1)
    for attempt in range(2):
        status = "Optimal"
        try:
            res = solver.solve_resident(rp)
            it = res.iteration()
        except Exception as e:  # noqa: BLE001
            status = type(e).__name__ + ": " + str(e)[:160]
            it = rp.last_iterations
        tr = rp.trace()
        sums = {}
        for name in ("M", "y", "b", "rP", "dy", "W", "t"):
            a = rp.debug_read(name)
            sums[name] = hashlib.sha1(a.tobytes()).hexdigest()[:12]
        Mdiag = rp.debug_read("M").reshape(m, -1)
        dg = np.diagonal(Mdiag[:, :m])
        info = dict(rank=rank, status=status, it=it, nrows=len(tr), sums=sums,
                    diag_min=float(np.nanmin(dg)), diag_nan=int(np.isnan(dg).sum()))
        allinfo = [None] * world
        dist.all_gather_object(allinfo, (info, tr.tolist()))
        if rank == 0:
            print("attempt", attempt)
            for inf, _ in allinfo:
                print(inf)
            t0 = np.array(allinfo[0][1])
            for r in range(1, world):
                tr_r = np.array(allinfo[r][1])
                k = min(len(t0), len(tr_r))
                neq = [i for i in range(k) if not np.array_equal(t0[i], tr_r[i])]
                print("rank", r, "rows", len(tr_r), "first differing iteration:", neq[:3])
                for i in neq[:2]:
                    print("  r0", t0[i])
                    print("  r%d" % r, tr_r[i])
            for i, row in enumerate(t0):
                print(i + 1, " ".join("%.3e" % v for v in row))
        dist.barrier()
    rp.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
