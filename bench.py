#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native interior-point path.

    python bench.py --gpus N --steps K --warmup W [--workload C1|C2|C3] [--impl reference]

A "step" is ONE COMPLETE SOLVE (blind start -> Optimal) of the workload LP, i.e. one pass of the
hot path (`InteriorPoint::solve`, /root/reference/src/solvers/interior_point/mod.rs:161-240) over
one synthetic problem.  metric = IPM iterations per second (BASELINE.json: "time-to-solve & IPM
iterations/s"); ms_per_step is the time-to-solve.

  value : iterations/s with the problem already resident in HBM (lpb_solve on a live context),
          timed with CUDA events on the stream the kernels are launched on, max over ranks.
  e2e   : the same through the public reference-shaped API (`InteriorPoint.solve(problem)`):
          context creation, H2D of A/b/c from pinned host memory, solve, D2H of x, teardown.
  roofline : the dominant kernel (K1, DMMA SYRK): m(m+1)n algorithmic flop per launch / its mean
          launch duration (CUDA events inside the library, on the launching stream).
  cpu_baseline : the oracle (NumPy/OpenBLAS restatement of the reference) on the host cores, on a
          bounded sample of the same workload.

N > 1: strong scaling -- the same LP with A column-sharded over the ranks, NCCL
all-reduce of M and of the A.w products.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {  # slack-form (m, n); BASELINE.json configs[0..2]
    "C1": (512, 1024),
    "C2": (4096, 8192),
    "C3": (16384, 32768),
}
DEFAULT_WORKLOAD = "C3"
NOMINAL_FP64_TFLOPS = 40.0  # B200 FP64 (tensor == vector), NVIDIA HGX B200 spec sheet


def synthetic_lp(m, n, seed):
    """SURVEY.md 8(d) generator (same bits as oracle.ipm_oracle.synthetic_lp; duplicated here so the
    product arm never imports the oracle)."""
    mh = m // 2
    n0 = n - mh
    rng = np.random.default_rng(seed)
    A0 = rng.standard_normal((m, n0))
    x0 = rng.uniform(0.5, 1.5, n0)
    s0 = rng.uniform(0.5, 1.5, mh)
    b = A0.dot(x0)
    b[:mh] += s0
    y0 = rng.standard_normal(m)
    y0[:mh] = -np.abs(y0[:mh])
    z0 = rng.uniform(0.5, 1.5, n0)
    c = A0.T.dot(y0) + z0
    return c, A0[:mh], b[:mh], A0[mh:], b[mh:]


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock + throttle reasons sampled every 100 ms during the timed region (NVML)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# --------------------------------------------------------------------------- CPU baseline / reference arm
def cpu_sample(m, n, seed, budget_s=25.0):
    """Time the oracle (the reference's algorithm on host cores, OpenBLAS threads) on a bounded
    sample of the workload.  Returns (iterations_per_s, sample_description, cores)."""
    from oracle import ipm_oracle as o
    cores = os.cpu_count() or 1
    args = synthetic_lp(m, n, seed)
    pb = o.build_problem(*args)
    del args
    if 2.0 * m * m * n < 2e11:  # small enough: time whole iterations of the real loop
        t0 = time.perf_counter()
        its = 0
        tr = []
        try:
            o.InteriorPoint().solve(pb, trace=tr, stop_after=1)
        except o.LinearProgramError:
            pass
        t_one = time.perf_counter() - t0
        k = int(max(1, min(25, budget_s / max(t_one, 1e-3))))
        t0 = time.perf_counter()
        tr = []
        try:
            o.InteriorPoint().solve(pb, trace=tr, stop_after=k)
        except o.LinearProgramError:
            pass
        its = max(1, len(tr))
        dt = time.perf_counter() - t0
        return its / dt, "oracle loop, first %d iterations of the %dx%d solve (incl. blind-start residuals)" % (
            its, m, n), cores
    # large: extrapolate one iteration from slices of its dominant pieces
    from scipy.linalg import lapack
    pt = o.blind_start(pb)
    Dinv = pt.x / pt.z
    A = pb.A
    r = 256
    while True:  # rows slice of M = A (Dinv * A^T): the reference runs the FULL GEMM (2 m^2 n flop)
        t0 = time.perf_counter()
        _ = A[:r].dot(Dinv[:, None] * A.T) if r >= m else A[:r].dot((A * Dinv).T)
        t_slice = time.perf_counter() - t0
        if t_slice > budget_s * 0.3 or r >= m:
            break
        r = min(m, r * 2)
    t_gemm = t_slice * m / r
    ms = min(m, 8192)
    rng = np.random.default_rng(1)
    B = rng.standard_normal((ms, ms + 16))
    S = B.dot(B.T) + ms * np.eye(ms)
    t0 = time.perf_counter()
    cfac, info = lapack.dpotrf(S, lower=0)
    t_potrf = (time.perf_counter() - t0) * (m / ms) ** 3
    rhs = rng.standard_normal(ms)
    t0 = time.perf_counter()
    lapack.dpotrs(cfac, rhs, lower=0)
    t_potrs = (time.perf_counter() - t0) * (m / ms) ** 2
    t0 = time.perf_counter()
    A.dot(pt.x)
    A.T.dot(pt.y)
    t_gemv2 = time.perf_counter() - t0
    t_iter = t_gemm + t_potrf + 4 * t_potrs + 6 * t_gemv2  # reference: 12 sweeps + 4 potrs per iteration
    desc = ("one %dx%d iteration extrapolated from slices: %d/%d rows of the A.D.A^T GEMM (%.1fs), dpotrf+dpotrs at "
            "%d scaled cubically/quadratically, 2 of the 12 GEMV sweeps measured in full" % (m, n, r, m, t_slice, ms))
    return 1.0 / t_iter, desc, cores


def run_reference_arm(args, m, n):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    desc, cores = "", 1
    for i in range(args.warmup + args.steps):
        budget = 20.0 if i >= args.warmup else 5.0
        v, desc, cores = cpu_sample(m, n, args.seed, budget_s=budget)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "ipm_iterations_per_s", "value": value, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "%s dense LP slack-form m=%d n=%d seed=%d" % (args.workload, m, n, args.seed)},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- product arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    m, n = WORKLOADS[args.workload]

    if args.impl == "reference":
        run_reference_arm(args, m, n)
        return

    import torch
    import lp_b200
    from lp_b200 import _ffi
    from lp_b200.api import ResidentProblem, ShardedProblem

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d" % args.gpus)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))

    lib = _ffi.load()
    stream = torch.cuda.current_stream().cuda_stream

    # ---- build the workload on the host (identical bits on every rank)
    t0 = time.perf_counter()
    c, A_ub, b_ub, A_eq, b_eq = synthetic_lp(m, n, args.seed)
    problem = lp_b200.Problem.target(c).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
    del A_ub, A_eq
    gen_s = time.perf_counter() - t0
    solver = lp_b200.InteriorPoint.default()

    if world > 1:
        rp = ShardedProblem(problem, rank, world, dist, stream=stream)
    else:
        rp = ResidentProblem(problem, stream=stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up
    iters = 0
    for _ in range(args.warmup):
        res = solver.solve_resident(rp)
        iters = res.iteration()
    barrier()

    # ---- timed region: EXACTLY K solves, CUDA events on the launching stream
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_sum = {}
    launches = 0
    total_iters = 0
    sampler.start()
    barrier()
    e0.record()
    for _ in range(args.steps):
        res = solver.solve_resident(rp)
        total_iters += res.iteration()
        p = rp.profile()
        launches += p["launches"]
        for k, v in p.items():
            prof_sum[k] = prof_sum.get(k, 0) + v
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = total_iters / (ms * 1e-3)
    fun = res.fun()

    # ---- e2e through the public API (host buffers, H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e and world == 1:
        rp.close()
        torch.cuda.synchronize()
        solver.solve(problem)  # warm (allocator, pinned pages)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        it_e2e = 0
        for _ in range(args.steps):
            r2 = solver.solve(problem)
            it_e2e += r2.iteration()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        e2e = {"value": it_e2e / dt, "unit": "iterations/s", "h2d_bytes_per_step": int((m * n + m + n) * 8),
               "d2h_bytes_per_step": int(n * 8 + 16), "ms_per_step": dt / args.steps * 1e3}
    elif world > 1:
        # sharded: the public call is ShardedProblem(...) + solve_resident; time upload + solve + download
        rp.close()
        barrier()
        t0 = time.perf_counter()
        it_e2e = 0
        for _ in range(args.steps):
            with ShardedProblem(problem, rank, world, dist, stream=stream) as sp:
                it_e2e += solver.solve_resident(sp).iteration()
        barrier()
        dt = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": it_e2e / dt, "unit": "iterations/s",
               "h2d_bytes_per_step": int((m * (n // world) + m + n // world) * 8),
               "d2h_bytes_per_step": int((n // world) * 8 + 16), "ms_per_step": dt / args.steps * 1e3}

    if rank == 0:
        # ---- roofline of the dominant kernel (K1 DMMA SYRK) and of the SYRK+Cholesky phase
        n_loc = n // world if world > 1 else n
        f_syrk = float(m) * (m + 1) * n_loc
        f_chol = float(m) ** 3 / 3.0
        syrk_launches = max(1, prof_sum.get("syrk_launches", 1))
        syrk_ms = prof_sum.get("syrk_ms", 0.0) / syrk_launches
        potrf_ms = prof_sum.get("potrf_ms", 0.0) / max(1, prof_sum.get("potrf_launches", 1))
        achieved = f_syrk / (syrk_ms * 1e-3) * 1e-12 if syrk_ms > 0 else None
        roofline = {
            "bound": "tensor", "kernel": "syrk_dmma_kernel (K1, A.diag(x/z).A^T, FP64 DMMA)",
            "achieved": achieved, "peak": NOMINAL_FP64_TFLOPS, "unit": "TFLOP/s",
            "frac": (achieved / NOMINAL_FP64_TFLOPS) if achieved else None, "traffic": None,
            "peak_source": "nominal B200 FP64 (MEASURED_PEAKS.json has no FP64 entry; measured DMMA/cuBLAS "
                           "DGEMM rates are in profiles/fp64_peaks_r01.json)",
            "flop_per_launch": f_syrk, "ms_per_launch": syrk_ms,
            "phase_syrk_plus_cholesky_tflops": ((f_syrk + f_chol) / ((syrk_ms + potrf_ms) * 1e-3) * 1e-12
                                                if syrk_ms + potrf_ms > 0 else None),
            "potrf_ms_per_launch": potrf_ms,
            "potrf_tflops": (f_chol / (potrf_ms * 1e-3) * 1e-12) if potrf_ms > 0 else None,
        }
        steps = args.steps
        phases = {k: prof_sum.get(k, 0.0) / steps for k in
                  ("total_ms", "syrk_ms", "potrf_ms", "solve_ms", "sweep_ms", "vector_ms", "comm_ms")}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            v, desc, cores = cpu_sample(m, n, args.seed)
            cpu = {"value": v, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": desc}
        line = {
            "metric": "ipm_iterations_per_s", "value": value, "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s dense LP slack-form m=%d n=%d seed=%d (ub+eq, SURVEY 8d generator)" % (
                args.workload, m, n, args.seed), "iterations_per_solve": total_iters / args.steps,
                "objective": fun, "l2": "inputs larger than L2 (A = %.0f MB)" % (m * n * 8 / 1e6),
                "parallelism": "1 GPU" if world == 1 else "A column-sharded over %d GPUs, NCCL all-reduce of M" % world,
                "host_generation_s": gen_s},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "phases_ms_per_solve": phases, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
