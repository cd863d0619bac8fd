#!/bin/bash
# first GPU call: FP64 peaks, parity tests, first bench lines, launch list
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > gpurun_out/gpu_info.txt 2>&1

echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=6 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== bench C2"; timeout 600 python bench.py --workload C2 --steps 2 --warmup 1 > gpurun_out/bench_C2.log 2>&1; echo "rc=$?"; tail -c 3000 gpurun_out/bench_C2.log
echo "== bench C3"; timeout 900 python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_C3.log 2>&1; echo "rc=$?"; tail -c 3000 gpurun_out/bench_C3.log
