"""Determinism stress of K2 (k_potrf) at large m:  python tools/stress_potrf.py M REPS cfg [cfg ...]
cfg = trsm_impl,update_impl,sync_each_launch.  Factors the same SPD matrix REPS times per config and
counts distinct bit patterns of L; locates the first differing entry when a repetition deviates."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from lp_b200 import _ffi
    m = int(sys.argv[1])
    reps = int(sys.argv[2])
    cfgs = [tuple(int(v) for v in c.split(",")) for c in sys.argv[3:]] or [(0, 0, 0)]
    lib = _ffi.load()
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(m, m + 256, dtype=torch.float64, device="cuda", generator=g)
    M0 = A @ A.T
    del A
    M0 += 0.05 * m * torch.eye(m, dtype=torch.float64, device="cuda")
    h = C.c_void_p()
    assert lib.lpb_create_bare(C.byref(h), m, m, None) == 0
    M = torch.empty_like(M0)
    for (ti, ui, sy) in cfgs:
        for key, val in ((b"trsm_impl", ti), (b"update_impl", ui), (b"sync_each_launch", sy)):
            assert lib.lpb_set_option(h, key, val) == 0
        ref = None
        counts = {}
        ms_tot = 0.0
        for r in range(reps):
            M.copy_(M0)
            info = C.c_int32(-1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            rc = lib.lpb_k_potrf(h, m, M.data_ptr(), m, C.byref(info))
            e1.record()
            torch.cuda.synchronize()
            assert rc == 0, _ffi.last_error()
            ms_tot += e0.elapsed_time(e1)
            L = torch.tril(M)
            s = int(L.view(torch.int64).sum().item())
            counts[s] = counts.get(s, 0) + 1
            if ref is None:
                ref = L.clone()
                ref_s = s
            elif s != ref_s:
                d = (L != ref)
                cols = d.any(dim=0).nonzero()
                rows = d.any(dim=1).nonzero()
                c0, r0 = int(cols[0]), int(rows[0])
                # the differing entries inside the first affected 128 x 128 tile
                tr, tc = r0 // 128 * 128, c0 // 128 * 128
                sub = d[tr:tr + 128, tc:tc + 128].nonzero()
                print("  cfg=%s rep=%d DIFFERS: n=%d first col=%d row=%d info=%d; in tile (%d,%d): %d entries, rows %s cols %s" % (
                    (ti, ui, sy), r, int(d.sum()), c0, r0, info.value, tr // 128, tc // 128, len(sub),
                    sorted(set(int(v) for v in sub[:, 0]))[:12], sorted(set(int(v) for v in sub[:, 1]))[:12]), flush=True)
            del L
        print("cfg trsm=%d update=%d sync=%d: %d reps, %d distinct results %s, %.2f ms avg" % (
            ti, ui, sy, reps, len(counts), sorted(counts.values(), reverse=True), ms_tot / reps), flush=True)
    lib.lpb_destroy(h)


if __name__ == "__main__":
    main()
