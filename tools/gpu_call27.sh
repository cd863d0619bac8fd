#!/bin/bash
# 1-GPU call: banded SYRK tile order (tests, C3 number, ncu --set full for the DRAM traffic), batched e2e diagnostic
set -u
mkdir -p gpurun_out
echo "== pytest gpu (kernels, solve)"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py -m gpu -q --tb=short --maxfail=10 > gpurun_out/pytest_gpu27.log 2>&1; rc=$?; echo "rc=$rc"; tail -4 gpurun_out/pytest_gpu27.log
echo "== batched e2e timing"; timeout 300 python tools/time_batched_e2e.py > gpurun_out/time_batched27.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/time_batched27.log
echo "== bench C3"; timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_C3_r01_v14.log 2>&1; echo "rc=$?"; tail -c 1300 gpurun_out/bench_C3_r01_v14.log
CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
echo "== ncu full K1 (C3)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:syrk_dmma_kernel<\(int\)0' -s 1 -c 1 -f -o gpurun_out/syrk_C3_r01_v14 $CMD > gpurun_out/ncu_full27.log 2>&1
echo "rc=$?"; grep -E "PROF|WARN|ERR" gpurun_out/ncu_full27.log | tail -2 | cut -c1-200
