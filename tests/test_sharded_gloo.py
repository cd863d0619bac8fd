"""N > 1 path on the CPU: world_size-2 gloo processes run the column-sharded oracle (the same
decomposition and the same collectives the GPU path uses) and must reproduce the
unsharded oracle; plus the host-side shard bookkeeping of lp_b200.api (column split, unique-id
broadcast, gather of x)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from lp_b200.api import shard_columns
    from oracle import ipm_oracle as o
    from oracle.sharded_oracle import TorchComm, solve_sharded
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    try:
        out = {}
        for (m, n, seed) in [(32, 80, 0), (64, 128, 3)]:
            pb = o.build_problem(*o.synthetic_lp(m, n, seed))
            shards = shard_columns(n, world, m // 2)
            c0, nk = shards[rank]
            st, xk, it = solve_sharded(np.ascontiguousarray(pb.A[:, c0:c0 + nk]), pb.b, pb.c[c0:c0 + nk], n,
                                       TorchComm(dist))
            # gather x the way ShardedProblem.gather_x does
            import torch
            per = max(nl for _, nl in shards)
            t = torch.zeros(per, dtype=torch.float64)
            t[:nk] = torch.from_numpy(xk)
            outs = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(outs, t)
            x = np.concatenate([o_.numpy()[:nl] for o_, (_, nl) in zip(outs, shards)])
            out[(m, n, seed)] = (st, x, it)
        # unique-id style broadcast: rank 0's 128 bytes must arrive everywhere
        payload = torch.arange(128, dtype=torch.uint8) if rank == 0 else torch.zeros(128, dtype=torch.uint8)
        dist.broadcast(payload, src=0)
        out["uid_ok"] = bool((payload == torch.arange(128, dtype=torch.uint8)).all())
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_column_sharded_iteration_matches_unsharded_oracle():
    import torch.multiprocessing as mp
    from oracle import ipm_oracle as o
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for key in [(32, 80, 0), (64, 128, 3)]:
        m, n, seed = key
        ref = o.InteriorPoint().solve(o.build_problem(*o.synthetic_lp(m, n, seed)))
        for rank in (0, 1):
            st, x, it = results[rank][key]
            assert st == "Optimal"
            assert abs(it - ref.iteration) <= 1
            assert np.abs(x[: len(ref.x)] - ref.x).max() < 1e-6
        np.testing.assert_array_equal(results[0][key][1], results[1][key][1])  # all ranks agree bit for bit
    assert results[0]["uid_ok"] and results[1]["uid_ok"]


def test_shard_columns_partition():
    from lp_b200.api import shard_columns
    for n in (1, 7, 128, 1001, 32768, 131072):
        for world in (1, 2, 3, 4, 8):
            for n_slack in (0, n // 4, n // 8 + 1, n):
                sh = shard_columns(n, world, n_slack)
                assert len(sh) == world
                assert sum(nl for _, nl in sh) == n
                pos = 0
                for c0, nl in sh:
                    assert c0 == pos or nl == 0
                    assert c0 % 2 == 0 or nl == 0   # even offsets keep 16-byte alignment of every shard
                    pos += nl
                # the dense columns (everything but the trailing slack block) are balanced to within one pair
                n_dense = n - n_slack if n_slack < n else n
                dense = [max(0, min(c0 + nl, n_dense) - c0) for c0, nl in sh]
                assert sum(dense) == n_dense
                assert max(dense) <= -(-n_dense // world) + 1


def test_local_comm_equals_unsharded():
    from oracle import ipm_oracle as o
    from oracle.sharded_oracle import LocalComm, solve_sharded
    pb = o.build_problem(*o.synthetic_lp(48, 100, 1))
    ref = o.InteriorPoint().solve(pb)
    st, x, it = solve_sharded(pb.A, pb.b, pb.c, pb.A.shape[1], LocalComm())
    assert st == "Optimal" and it == ref.iteration
    assert np.abs(x[: len(ref.x)] - ref.x).max() < 1e-9
