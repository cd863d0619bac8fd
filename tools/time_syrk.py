"""Time K1 (k_syrk_adat) alone for every accumulation variant, CUDA events on the launching stream, and measure what
each variant's summation order costs in accuracy.   python tools/time_syrk.py [m n_dense]   (default: the C3 shape)

Variants: "syrk_chain" = 1 (one register chain over the whole K extent, the round-1 kernel) and the blocked
accumulation with "syrk_flush_blocks" = 32 ... 4096 K-blocks of 16 columns between two flushes (a period longer
than the K extent never flushes: it times the flush CODE without its memory traffic).
Accuracy: (a) max |M - M_ref| / (|A| D |A|^T) over the lower triangle, M_ref summed by cuBLAS in blocks of 1024
columns; (b) ulp error of the first 512 diagonal entries (n same-sign terms each) against extended precision."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from lp_b200 import _ffi
    lib = _ffi.load()
    m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16384, 24576)
    g = torch.Generator("cuda").manual_seed(1)
    A = torch.randn((m, n), dtype=torch.float64, device="cuda", generator=g)
    d = torch.exp(4.0 * torch.randn(n, dtype=torch.float64, device="cuda", generator=g))
    Mref = torch.zeros((m, m), dtype=torch.float64, device="cuda")
    for k0 in range(0, n, 1024):
        Mref += (A[:, k0:k0 + 1024] * d[k0:k0 + 1024]) @ A[:, k0:k0 + 1024].T
    Mabs = (A.abs() * d) @ A.abs().T
    low = torch.tril(torch.ones((m, m), dtype=torch.bool, device="cuda"))
    rows = min(m, 512)
    Ah = A[:rows].cpu().numpy().astype(np.longdouble)
    exact = ((Ah * Ah) * d.cpu().numpy().astype(np.longdouble)).sum(axis=1)
    del Ah
    ulp = np.spacing(exact.astype(np.float64))
    stream = torch.cuda.current_stream().cuda_stream
    h = C.c_void_p()
    assert lib.lpb_create_bare(C.byref(h), m, n, C.c_void_p(stream)) == 0
    M = torch.empty((m, m), dtype=torch.float64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    variants = [("chain", 1, 128)] + [("flush%d" % fb, 0, fb) for fb in (32, 64, 128, 256, 512, 4096)]
    for name, chain, fb in variants:
        assert lib.lpb_set_option(h, b"syrk_chain", chain) == 0
        assert lib.lpb_set_option(h, b"syrk_flush_blocks", fb) == 0
        ts = []
        for rep in range(4):
            M.fill_(float("nan"))
            torch.cuda.synchronize()
            e0.record()
            assert lib.lpb_k_syrk_adat(h, m, n, A.data_ptr(), n, d.data_ptr(), M.data_ptr(), m) == 0
            e1.record()
            torch.cuda.synchronize()
            if rep:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        err = ((M - Mref).abs() / Mabs)[low].max().item()
        dg = torch.diagonal(M)[:rows].cpu().numpy().astype(np.longdouble)
        u = (np.abs(dg - exact) / ulp).astype(np.float64)
        rec = dict(ms=ts[len(ts) // 2], ms_best=ts[0], tflops=m * (m + 1.0) * n / (ts[len(ts) // 2] * 1e-3) * 1e-12,
                   max_rel_err=err, diag_ulp_rms=float(np.sqrt((u ** 2).mean())), diag_ulp_max=float(u.max()))
        out[name] = rec
        print("%-10s %s" % (name, json.dumps(rec)), flush=True)
    # the checker against itself: one cuBLAS call (a single chain per entry) vs the K-blocked sum
    Mone = (A * d) @ A.T
    dg = torch.diagonal(Mone)[:rows].cpu().numpy().astype(np.longdouble)
    u = (np.abs(dg - exact) / ulp).astype(np.float64)
    print("cublas-one-call  max_rel_err %.3e  diag_ulp_rms %.2f max %.1f" % (
        ((Mone - Mref).abs() / Mabs)[low].max().item(), float(np.sqrt((u ** 2).mean())), float(u.max())))
    dg = torch.diagonal(Mref)[:rows].cpu().numpy().astype(np.longdouble)
    u = (np.abs(dg - exact) / ulp).astype(np.float64)
    print("cublas-blocked   diag_ulp_rms %.2f max %.1f" % (float(np.sqrt((u ** 2).mean())), float(u.max())))
    lib.lpb_destroy(h)


if __name__ == "__main__":
    main()
