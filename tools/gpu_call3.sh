#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_gpu.log
CMD="python bench.py --workload C2 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
echo "== plain + ncu launch list (C2)"
$CMD > gpurun_out/plain_C2.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_C2.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_launches.log; wc -l gpurun_out/launches_C2.csv
echo "== ncu full on syrk (C2)"
$CMD > gpurun_out/plain_C2b.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:syrk_dmma -s 1 -c 2 -o gpurun_out/syrk_C2 $CMD > gpurun_out/ncu_full.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
