"""Time K2 (k_potrf) and K3 (k_potrs, every solve_impl) alone, CUDA events on the launching stream, against
cuSOLVER potrf / potrs on the same matrix (checkers only).   python tools/time_kernels.py 4096 16384"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(torch, fn, reps, setup=None):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        if setup:
            setup()
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    import torch
    from lp_b200 import _ffi
    lib = _ffi.load()
    out = {}
    stream = torch.cuda.current_stream().cuda_stream
    for m in [int(a) for a in sys.argv[1:]]:
        g = torch.Generator("cuda").manual_seed(m)
        Bm = torch.randn((m, m + 8), dtype=torch.float64, device="cuda", generator=g)
        M = Bm @ Bm.T + 0.5 * torch.eye(m, dtype=torch.float64, device="cuda")
        del Bm
        W = torch.empty_like(M)
        h = C.c_void_p()
        assert lib.lpb_create_bare(C.byref(h), m, m, C.c_void_p(stream)) == 0
        info = C.c_int32(-1)
        rec = {}
        for look in (1, 0):
            assert lib.lpb_set_option(h, b"potrf_lookahead", look) == 0
            med, best = timed(torch, lambda: lib.lpb_k_potrf(h, m, W.data_ptr(), m, C.byref(info)), 5,
                              setup=lambda: W.copy_(M))
            rec["potrf_lookahead%d_ms" % look] = med
            rec["potrf_lookahead%d_tflops" % look] = m ** 3 / 3.0 / (med * 1e-3) * 1e-12
        assert info.value == 0
        med, _ = timed(torch, lambda: torch.linalg.cholesky_ex(M), 5)
        rec["cusolver_potrf_ms"] = med
        assert lib.lpb_set_option(h, b"potrf_lookahead", 1) == 0
        W.copy_(M)
        assert lib.lpb_k_potrf(h, m, W.data_ptr(), m, C.byref(info)) == 0
        rhs = torch.randn((2, m), dtype=torch.float64, device="cuda", generator=g)
        X = torch.empty_like(rhs)
        for impl in (0, 3, 1):
            assert lib.lpb_set_option(h, b"solve_impl", impl) == 0
            for nrhs in (1, 2):
                med, best = timed(torch, lambda: lib.lpb_k_potrs(h, m, W.data_ptr(), m, X.data_ptr(), nrhs), 7,
                                  setup=lambda: X.copy_(rhs))
                rec["potrs_impl%d_nrhs%d_ms" % (impl, nrhs)] = med
                rec["potrs_impl%d_nrhs%d_gbs" % (impl, nrhs)] = 8.0 * m * m / (med * 1e-3) / 1e9  # L read twice: 2 x 4 m^2 bytes
        Lt = torch.tril(W)
        med, _ = timed(torch, lambda: torch.cholesky_solve(rhs.T.contiguous(), Lt), 5)
        rec["cusolver_potrs_2rhs_ms"] = med
        lib.lpb_destroy(h)
        out[str(m)] = rec
        print(m, json.dumps(rec), flush=True)
        del M, W, Lt
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    main()
