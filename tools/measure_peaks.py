"""Step 0: FP64 denominators on this B200 (run under gpurun).  Writes gpurun_out/fp64_peaks.json.

  * hand-written DFMA / DMMA issue-rate microbenchmarks (tools/fp64_peaks.cu)
  * library baselines, NOT linked into the product: cuBLAS DGEMM (torch.matmul fp64) and
    cuSOLVER DPOTRF (torch.linalg.cholesky fp64)
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def build_tool():
    src = os.path.join(ROOT, "tools", "fp64_peaks.cu")
    exe = os.path.join(ROOT, "tools", "_build", "fp64_peaks")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(src):
        subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3",
                               "-lineinfo", src, "-o", exe])
    return exe


def main():
    os.makedirs(OUT, exist_ok=True)
    res = {}
    exe = build_tool()
    if "--build-only" in sys.argv:
        return
    res["micro"] = json.loads(subprocess.check_output([exe]).decode())
    import torch
    dev = "cuda"

    def tm(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    lib = {}
    for n in (4096, 8192):
        a = torch.randn(n, n, dtype=torch.float64, device=dev)
        b = torch.randn(n, n, dtype=torch.float64, device=dev)
        ms = tm(lambda: torch.matmul(a, b))
        lib["cublas_dgemm_%d_tflops" % n] = 2.0 * n ** 3 / ms * 1e-9
        del a, b
    # A (m x n) times A^T: the SYRK-shaped product as a full DGEMM (what the reference executes)
    m, n = 4096, 8192
    a = torch.randn(m, n, dtype=torch.float64, device=dev)
    ms = tm(lambda: torch.matmul(a, a.t()))
    lib["cublas_dgemm_AAt_4096x8192_tflops_full"] = 2.0 * m * m * n / ms * 1e-9
    lib["cublas_dgemm_AAt_4096x8192_ms"] = ms
    del a
    for n in (4096, 16384):
        b = torch.randn(n, n + 64, dtype=torch.float64, device=dev)
        spd = b @ b.t() + n * torch.eye(n, dtype=torch.float64, device=dev)
        del b
        ms = tm(lambda: torch.linalg.cholesky(spd), reps=2)
        lib["cusolver_dpotrf_%d_tflops" % n] = n ** 3 / 3.0 / ms * 1e-9
        lib["cusolver_dpotrf_%d_ms" % n] = ms
        del spd
    res["library"] = lib
    res["when"] = time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())
    with open(os.path.join(OUT, "fp64_peaks.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
