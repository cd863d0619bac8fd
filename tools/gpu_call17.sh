#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/trace_solve.py C3 > gpurun_out/trace_C3_h.log 2>&1; echo "rc=$?"; head -2 gpurun_out/trace_C3_h.log; grep -E "^(1[6-9]|2[0-9]) gpu|gpu kappa|ora kappa" gpurun_out/trace_C3_h.log | sed -n '31,60p' | cut -c1-250
timeout 300 python tools/trace_solve.py C2 > gpurun_out/trace_C2_h.log 2>&1; echo "rc=$?"; head -2 gpurun_out/trace_C2_h.log; tail -9 gpurun_out/trace_C2_h.log | cut -c1-250
echo "== pytest solve"; timeout 900 python -m pytest tests/test_gpu_solve.py tests/test_cpp_host.py tests/test_gpu_batched.py -m gpu -q --tb=short --maxfail=10 > gpurun_out/pytest_gpu17.log 2>&1; rc=$?; echo "rc=$rc"; tail -5 gpurun_out/pytest_gpu17.log
echo "== bench C3"; timeout 500 python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_C3_v6.log 2>&1; echo "rc=$?"; tail -c 700 gpurun_out/bench_C3_v6.log
