#!/bin/bash
set -u
mkdir -p gpurun_out
for v in "0 4 1" "0 0 1" "0 5 1"; do
echo "== in-solve verify C3 variant $v"; timeout 400 python tools/diag_potrf_insolve.py 16384 32768 $v > gpurun_out/insolve10_$(echo $v | tr -d ' ').log 2>&1; echo "rc=$?"; tail -12 gpurun_out/insolve10_$(echo $v | tr -d ' ').log | cut -c1-200
done
echo "== pytest kernels"; timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --tb=short -x > gpurun_out/pytest_gpu10.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu10.log
echo "== bench C3"; timeout 500 python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_C3_v4.log 2>&1; echo "rc=$?"; tail -c 1600 gpurun_out/bench_C3_v4.log
