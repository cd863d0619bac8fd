"""Is the FP64 tensor-core accumulation (mma.sync.m8n8k4.f64 -> DMMA) rounded to nearest, or biased?

    python tools/dmma_bias.py

M = A A^T with A made of 21-bit dyadic rationals, so every product a_ik a_jk is EXACT in FP64 and the only rounding is
the accumulation; the exact sums come from Python integers.  Reports the SIGNED error of each accumulation scheme in
units of ulp(result): an unbiased scheme has mean ~ 0 and std ~ sqrt(#roundings) / sqrt(12); a truncating one has a
mean of about -#roundings / 2.   Schemes: the DMMA kernel (K1), the plain-DFMA kernel (syrk_impl = 1), cuBLAS DGEMM
(torch.matmul) and NumPy/OpenBLAS on the host.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from lp_b200 import _ffi
    lib = _ffi.load()
    m, n = 192, 8192
    rng = np.random.default_rng(0)
    for name, lo, hi in (("positive terms (a in [0.5, 1.5))", 2 ** 19, 3 * 2 ** 19), ("mixed signs", -2 ** 20, 2 ** 20)):
        Ai = rng.integers(lo, hi, size=(m, n), dtype=np.int64)
        A = Ai.astype(np.float64) * 2.0 ** -20            # exact
        exact_int = Ai.astype(object) @ Ai.astype(object).T   # Python integers: exact sums of exact products
        scale = 2.0 ** -40
        stream = torch.cuda.current_stream().cuda_stream
        dA = torch.from_numpy(A).cuda()
        out = {}
        for impl, label in ((0, "K1 DMMA kernel"), (1, "plain DFMA kernel")):
            h = C.c_void_p()
            assert lib.lpb_create_bare(C.byref(h), m, n, C.c_void_p(stream)) == 0
            assert lib.lpb_set_option(h, b"syrk_impl", impl) == 0
            dM = torch.zeros((m, m), dtype=torch.float64, device="cuda")
            assert lib.lpb_k_syrk_adat(h, m, n, dA.data_ptr(), n, None, dM.data_ptr(), m) == 0
            out[label] = dM.cpu().numpy()
            lib.lpb_destroy(h)
        out["cuBLAS DGEMM (torch.matmul)"] = (dA @ dA.T).cpu().numpy()
        out["NumPy / OpenBLAS (host)"] = A @ A.T
        print("%s, n = %d terms per entry, %d entries of the lower triangle:" % (name, n, m * (m + 1) // 2))
        low = np.tril_indices(m)
        for label, M in out.items():
            errs = []
            for i, j in zip(*low):
                ex = exact_int[i, j]
                got = int(round(M[i, j] / scale))        # M / 2^-40 is an integer-valued double (exact division by a power of 2)
                ulp = np.spacing(abs(M[i, j])) / scale
                errs.append((got - ex) / ulp)
            errs = np.array(errs, dtype=np.float64)
            print("   %-30s signed error / ulp: mean %+8.3f  std %7.3f  min %+8.2f  max %+8.2f" % (
                label, errs.mean(), errs.std(), errs.min(), errs.max()))


if __name__ == "__main__":
    main()
