#!/bin/bash
# 1-GPU call: diagonal-in-register pivot chain (potf2 + batched), pinned C4 e2e, ncu of the batched kernel
set -u
mkdir -p gpurun_out
echo "== pytest gpu (kernels, solve, batched)"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py tests/test_gpu_batched.py -m gpu -q --tb=short --maxfail=10 > gpurun_out/pytest_gpu25.log 2>&1; rc=$?; echo "rc=$rc"; tail -4 gpurun_out/pytest_gpu25.log
echo "== bench C4"; timeout 600 python bench.py --workload C4 --steps 3 --warmup 3 > gpurun_out/bench_C4_r01_v13.log 2>&1; echo "rc=$?"; tail -c 1900 gpurun_out/bench_C4_r01_v13.log
echo "== bench C2"; python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2_r01_v13.log 2>&1; echo "rc=$?"; tail -c 1400 gpurun_out/bench_C2_r01_v13.log
CMD="python bench.py --workload C4 --steps 1 --warmup 0 --no-cpu-baseline"
echo "== plain C4"; $CMD > gpurun_out/plain_C4_25.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_C4_25.log; }
echo "== ncu full batched (C4)"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:batched_ipm_kernel' -c 1 -f -o gpurun_out/batched_C4_r01_v13 $CMD > gpurun_out/ncu_full25.log 2>&1
echo "rc=$?"; grep -E "PROF|WARN|ERR" gpurun_out/ncu_full25.log | tail -2 | cut -c1-200
