#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest kernels+solve"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py tests/test_cpp_host.py -m gpu -q --tb=short --maxfail=10 > gpurun_out/pytest_gpu13.log 2>&1; rc=$?; echo "rc=$rc"; tail -25 gpurun_out/pytest_gpu13.log
timeout 300 python tools/trace_solve.py C3 > gpurun_out/trace_C3_f.log 2>&1; echo "rc=$?"; head -3 gpurun_out/trace_C3_f.log; tail -9 gpurun_out/trace_C3_f.log | cut -c1-220
echo "== bench C2"; timeout 300 python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_C2_v5.log 2>&1; echo "rc=$?"; tail -c 1500 gpurun_out/bench_C2_v5.log
echo "== bench C3"; timeout 500 python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_C3_v5.log 2>&1; echo "rc=$?"; tail -c 1500 gpurun_out/bench_C3_v5.log
