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

if "reference" in sys.argv[1:]:
    # The CPU arm uses every host thread -- also under torchrun, which exports OMP_NUM_THREADS=1 to its ranks.
    # Must happen before NumPy loads its BLAS.
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = str(os.cpu_count() or 1)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {  # slack-form (m, n); BASELINE.json configs[0..4]
    "C1": (512, 1024),
    "C2": (4096, 8192),
    "C3": (16384, 32768),
    "C4": (64, 128),        # batched: 8192 independent LPs, one CTA per problem, batch sharded over ranks
    "C5": (32768, 131072),  # A generated on the device per column shard (never on the host)
}
C4_BATCH = 8192
DEFAULT_WORKLOAD = "C3"
NOMINAL_FP64_TFLOPS = 40.0  # B200 FP64 (tensor == vector), NVIDIA HGX B200 spec sheet


def workload_label(name, m, n, seed):
    return "%s dense LP slack-form m=%d n=%d seed=%d (ub+eq, SURVEY 8d generator)" % (name, m, n, seed)


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
def cpu_problem(m, n, seed):
    """The workload LP as the oracle's Problem (built once per process: generation is not what is timed)."""
    from oracle import ipm_oracle as o
    return o.build_problem(*synthetic_lp(m, n, seed))


def cpu_sample(m, n, seed, budget_s=25.0, pb=None):
    """Time the oracle (the reference's algorithm on host cores, OpenBLAS threads) on a bounded
    sample of the workload.  Returns (iterations_per_s, sample_description, cores)."""
    from oracle import ipm_oracle as o
    cores = os.cpu_count() or 1
    if pb is None:
        pb = cpu_problem(m, n, seed)
    if 2.0 * m * m * n < 1e12:  # small enough (C1, C2): time whole iterations of the real loop
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
    t0 = time.perf_counter()
    B = Dinv[:, None] * A.T          # the reference's n x m temporary (newton_equations.rs:57), timed in full
    t_temp = time.perf_counter() - t0
    r = 256
    while True:  # row slice of M = A B: the reference runs the FULL GEMM (2 m^2 n flop)
        t0 = time.perf_counter()
        _ = A[:r].dot(B)
        t_slice = time.perf_counter() - t0
        if t_slice > budget_s * 0.3 or r >= m:
            break
        r = min(m, r * 2)
    del B
    t_gemm = t_temp + t_slice * m / r
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
    pb = cpu_problem(m, n, args.seed)
    for i in range(args.warmup + args.steps):
        budget = 20.0 if i >= args.warmup else 5.0
        v, desc, cores = cpu_sample(m, n, args.seed, budget_s=budget, pb=pb)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "ipm_iterations_per_s", "value": value, "unit": "iterations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_label(args.workload, m, n, args.seed)},
        "cpu_baseline": {"value": value, "unit": "iterations/s", "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- shared pieces of the product arm
def _sync_max_ms(torch, dist, ms):
    if dist is None:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _roofline(prof_sum, m, n_loc):
    """K1 roofline.  `achieved` counts the flop the kernel EXECUTES: m(m+1) * syrk_cols, where syrk_cols
    is n minus the trailing slack columns that liblpb200 folds into the diagonal of M instead of
    contracting over them (lpb_profile.syrk_cols).  SURVEY 8(d)'s algorithmic figure m(m+1)n -- what a
    structure-blind SYRK would execute -- divided by the same time is reported beside it."""
    launches = max(1, prof_sum.get("syrk_launches", 1))
    solves = max(1, prof_sum.get("solves", 1))
    n_exec = prof_sum.get("syrk_cols", 0) / solves or n_loc
    f_alg = float(m) * (m + 1) * n_loc
    f_syrk = float(m) * (m + 1) * n_exec
    f_chol = float(m) ** 3 / 3.0
    syrk_ms = prof_sum.get("syrk_ms", 0.0) / launches
    potrf_ms = prof_sum.get("potrf_ms", 0.0) / max(1, prof_sum.get("potrf_launches", 1))
    achieved = f_syrk / (syrk_ms * 1e-3) * 1e-12 if syrk_ms > 0 else None
    return {
        "bound": "tensor", "kernel": "syrk_dmma_kernel (K1, A.diag(x/z).A^T, FP64 DMMA)",
        "achieved": achieved, "peak": NOMINAL_FP64_TFLOPS, "unit": "TFLOP/s",
        "frac": (achieved / NOMINAL_FP64_TFLOPS) if achieved else None, "traffic": SYRK_TRAFFIC.get((m, int(n_exec))),
        "peak_source": "nominal B200 FP64 (MEASURED_PEAKS.json has no FP64 entry; measured on this pool: DMMA issue "
                       "peak 36.95, cuBLAS DGEMM 35.4 TFLOP/s, profiles/fp64_peaks_r01.json)",
        "flop_per_launch": f_syrk, "ms_per_launch": syrk_ms, "syrk_cols": n_exec,
        "algorithmic_flop_per_launch": f_alg,
        "algorithmic_tflops": (f_alg / (syrk_ms * 1e-3) * 1e-12) if syrk_ms > 0 else None,
        "phase_syrk_plus_cholesky_tflops": ((f_syrk + f_chol) / ((syrk_ms + potrf_ms) * 1e-3) * 1e-12
                                            if syrk_ms + potrf_ms > 0 else None),
        "potrf_ms_per_launch": potrf_ms,
        "potrf_tflops": (f_chol / (potrf_ms * 1e-3) * 1e-12) if potrf_ms > 0 else None,
    }


# dram__bytes_read.sum + dram__bytes_write.sum of ONE syrk_dmma_kernel launch, from the committed
# `ncu --set full` captures (profiles/); keyed by (m, n_local).
SYRK_TRAFFIC = {
    # C3 on 1 GPU, banded tile order: 46.05 GB read + 1.08 GB written (profiles/ncu_syrk_C3_r01_v14.txt);
    # algorithmic bytes: 3.22 GB of A (dense columns) + 1.07 GB of M.  The kernel is DMMA-bound (DRAM at 3 %
    # of peak); the re-reads are operand tiles streamed once per wave of 148 tiles.
    (16384, 24576): 46.048303e9 + 1.082822e9,
}


def _timed_solves(args, torch, dist, solver, rp, local_rank):
    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        solver.solve_resident(rp)
    barrier()
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof_sum, launches, total_iters, res = {}, 0, 0, None
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
        prof_sum["solves"] = prof_sum.get("solves", 0) + 1
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = _sync_max_ms(torch, dist, e0.elapsed_time(e1))
    return ms, total_iters, launches, prof_sum, clocks, res


def run_device_synthetic(args, torch, dist, rank, local_rank, world, stream, m, n):
    """C5 (and any size with --device-synthetic): A is generated per column shard on the device."""
    import lp_b200
    from lp_b200.api import SyntheticShardedProblem
    solver = lp_b200.InteriorPoint.default()
    t0 = time.perf_counter()
    rp = SyntheticShardedProblem(m, n, args.seed, rank, world, dist, stream=stream)
    rp.set_option("potrf_dist", args.potrf_dist)
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    ms, total_iters, launches, prof_sum, clocks, res = _timed_solves(args, torch, dist, solver, rp, local_rank)
    value = total_iters / (ms * 1e-3)
    rp.close()
    # e2e: context creation + on-device generation of the shard + solve + D2H / gather of x
    e2e = None
    if not args.no_e2e:
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        it_e2e = 0
        for _ in range(args.steps):
            with SyntheticShardedProblem(m, n, args.seed, rank, world, dist, stream=stream) as sp:
                it_e2e += solver.solve_resident(sp).iteration()
        torch.cuda.synchronize()
        dt = _sync_max_ms(torch, dist, (time.perf_counter() - t0) * 1e3) * 1e-3
        e2e = {"value": it_e2e / dt, "unit": "iterations/s", "h2d_bytes_per_step": 0,
               "d2h_bytes_per_step": int(rp.n * 8 + 16), "ms_per_step": dt / args.steps * 1e3,
               "note": "inputs are generated on the device (the %.1f GB matrix never exists on the host); "
                       "the timed region covers generation + solve + D2H of x" % (m * n * 8 / 1e9)}
    if rank == 0:
        n_loc = rp.n
        steps = args.steps
        phases = {k: prof_sum.get(k, 0.0) / steps for k in
                  ("total_ms", "syrk_ms", "potrf_ms", "solve_ms", "sweep_ms", "vector_ms", "comm_ms")}
        line = {
            "metric": "ipm_iterations_per_s", "value": value, "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s dense LP slack-form m=%d n=%d seed=%d, generated on the device per column "
                                   "shard (counter-based N(0,1), SURVEY 8d construction)" % (args.workload, m, n, args.seed),
                       "iterations_per_solve": total_iters / args.steps, "objective": res.fun(),
                       "l2": "inputs larger than L2 (A shard = %.0f MB)" % (m * n_loc * 8 / 1e6),
                       "parallelism": "1 GPU" if world == 1 else
                       "A column-sharded over %d GPUs, NCCL all-reduce of M, %s Cholesky" % (
                           world, "panel-broadcast distributed" if args.potrf_dist else "replicated"),
                       "device_generation_s": gen_s},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": _roofline(prof_sum, m, n_loc),
            "phases_ms_per_solve": phases, "cpu_baseline": None,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        _ffi_finalize()
        dist.destroy_process_group()


def run_batched(args, torch, dist, rank, local_rank, world, stream):
    """C4: `batch` independent 64x128 LPs, one CTA per problem; the batch is sharded over the ranks
    with no collective on the data path."""
    import lp_b200
    from lp_b200 import _ffi
    lib = _ffi.load()
    m, n = WORKLOADS["C4"]
    batch = args.batch
    per = -(-batch // world)
    lo, hi = min(batch, rank * per), min(batch, (rank + 1) * per)
    nb = hi - lo
    t0 = time.perf_counter()
    A = lp_b200.pinned_empty((nb, m, n))  # host inputs of the e2e leg live in pinned memory (H2D at PCIe rate)
    b = lp_b200.pinned_empty((nb, m))
    c = lp_b200.pinned_empty((nb, n))
    for i in range(nb):  # seeds 1000 + i
        cc, A_ub, b_ub, A_eq, b_eq = synthetic_lp(m, n, 1000 + lo + i)
        pb = lp_b200.Problem.target(cc).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
        A[i], b[i], c[i] = pb.A(), pb.b(), pb.c()
    gen_s = time.perf_counter() - t0
    solver = lp_b200.InteriorPoint.default()
    dA, db, dc = torch.from_numpy(A).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(c).cuda()
    dx = torch.zeros((nb, n), dtype=torch.float64, device="cuda")
    dfun = torch.zeros(nb, dtype=torch.float64, device="cuda")
    dit = torch.zeros(nb, dtype=torch.int64, device="cuda")
    dst = torch.zeros(nb, dtype=torch.int32, device="cuda")

    def resident():
        rc = lib.lpb_solve_batched(nb, m, n, dA.data_ptr(), db.data_ptr(), dc.data_ptr(), C.byref(solver._o),
                                   dx.data_ptr(), dfun.data_ptr(), dit.data_ptr(), dst.data_ptr(),
                                   _ffi.LPB_MEM_DEVICE, C.c_void_p(stream))
        assert rc == 0, _ffi.last_error()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        resident()
    barrier()
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    barrier()
    e0.record()
    for _ in range(args.steps):
        resident()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = _sync_max_ms(torch, dist, e0.elapsed_time(e1))
    its = dit.sum().to(torch.float64).reshape(1)
    n_ok = (dst == 0).sum().to(torch.float64).reshape(1)
    if dist is not None:
        dist.all_reduce(its)
        dist.all_reduce(n_ok)
    total_iters = float(its.item())
    value = total_iters * args.steps / (ms * 1e-3)
    # e2e: host arrays in, host arrays out through lp_b200.solve_batched (H2D of A, b, c + D2H of x, fun, ...)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = lp_b200.solve_batched(A, b, c, n_slack=m // 2, solver=solver, stream=stream)
    torch.cuda.synchronize()
    dt = _sync_max_ms(torch, dist, (time.perf_counter() - t0) * 1e3) * 1e-3
    e2e = {"value": total_iters * args.steps / dt, "unit": "iterations/s",
           "h2d_bytes_per_step": int((A.size + b.size + c.size) * 8), "d2h_bytes_per_step": int(nb * (n * 8 + 8 + 8 + 4)),
           "ms_per_step": dt / args.steps * 1e3}
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            from oracle import ipm_oracle as o
            k = min(nb, 256)
            t0 = time.perf_counter()
            it_cpu = 0
            for i in range(k):
                it_cpu += o.InteriorPoint().solve(o.Problem(A[i], b[i], c[i], 0.0, m // 2)).iteration
            dtc = time.perf_counter() - t0
            cpu = {"value": it_cpu / dtc, "unit": "iterations/s", "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": "oracle looped over the first %d of the %d LPs" % (k, batch)}
        # one-time load of each LP into shared memory is the only HBM traffic of the kernel
        bytes_per_lp = (m * n + m + n) * 8 + n * 8 + 24
        line = {
            "metric": "ipm_iterations_per_s", "value": value, "unit": "iterations/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C4 batched: %d independent dense LPs slack-form m=%d n=%d, seeds 1000+i, one CTA "
                                   "per problem" % (batch, m, n),
                       "lps_per_s": batch * args.steps / (ms * 1e-3), "optimal": int(n_ok.item()),
                       "iterations_per_lp": total_iters / batch,
                       "l2": "whole batch (%.0f MB) larger than L2" % (batch * bytes_per_lp / 1e6),
                       "parallelism": "batch sharded over %d GPU(s), no collective" % world,
                       "host_generation_s": gen_s},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(args.steps),
            "roofline": {"bound": "hbm", "kernel": "batched_ipm_kernel (K6): latency / shared-memory bound; HBM only "
                                                   "for the one-time load of each LP",
                         "achieved": batch * bytes_per_lp / (ms / args.steps * 1e-3) / 1e9 / world, "peak": _hbm_peak(),
                         "unit": "GB/s", "frac": batch * bytes_per_lp / (ms / args.steps * 1e-3) / 1e9 / world / _hbm_peak(),
                         "traffic": None},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        _ffi_finalize()
        dist.destroy_process_group()


def _ffi_finalize():
    from lp_b200 import _ffi
    _ffi.load().lpb_comm_finalize()


def _hbm_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0  # B200_PROFILING.md fallback


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
    ap.add_argument("--device-synthetic", action="store_true",
                    help="generate the LP on the device per column shard (always on for C5)")
    ap.add_argument("--batch", type=int, default=C4_BATCH)
    ap.add_argument("--potrf-dist", type=int, default=1, choices=[0, 1],
                    help="N > 1: 1 = distributed panel-broadcast Cholesky (default), 0 = replicated on every rank")
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

    if args.workload == "C4":
        run_batched(args, torch, dist, rank, local_rank, world, stream)
        return
    if args.workload == "C5" or args.device_synthetic:
        run_device_synthetic(args, torch, dist, rank, local_rank, world, stream, m, n)
        return

    # ---- build the workload on the host (identical bits on every rank)
    t0 = time.perf_counter()
    c, A_ub, b_ub, A_eq, b_eq = synthetic_lp(m, n, args.seed)
    problem = lp_b200.Problem.target(c).ub(A_ub, b_ub).eq(A_eq, b_eq).build()
    del A_ub, A_eq
    gen_s = time.perf_counter() - t0
    solver = lp_b200.InteriorPoint.default()

    if world > 1:
        rp = ShardedProblem(problem, rank, world, dist, stream=stream)
        rp.set_option("potrf_dist", args.potrf_dist)
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
        prof_sum["solves"] = prof_sum.get("solves", 0) + 1
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
    n_loc_rank = rp.n

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
        roofline = _roofline(prof_sum, m, n_loc_rank)
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
            "config": {"workload": workload_label(args.workload, m, n, args.seed),
                       "iterations_per_solve": total_iters / args.steps,
                "objective": fun, "l2": "inputs larger than L2 (A = %.0f MB)" % (m * n * 8 / 1e6),
                "parallelism": "1 GPU" if world == 1 else "A column-sharded over %d GPUs, NCCL all-reduce of M, %s Cholesky" % (
                           world, "panel-broadcast distributed" if args.potrf_dist else "replicated"),
                "host_generation_s": gen_s},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "phases_ms_per_solve": phases, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        _ffi_finalize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
