#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_gpu.log
echo "== bench C2"; timeout 600 python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2.log 2>&1; echo "rc=$?"; tail -c 2500 gpurun_out/bench_C2.log
echo "== bench C3"; timeout 900 python bench.py --workload C3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_C3.log 2>&1; echo "rc=$?"; tail -c 2500 gpurun_out/bench_C3.log
CMD="python bench.py --workload C2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
echo "== plain + ncu launch list (C2)"
$CMD > gpurun_out/plain_C2.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_C2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches_C2.csv
