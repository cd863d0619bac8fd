#!/bin/bash
# 1-GPU call: whole GPU suite, default bench line (C3, e2e + cpu_baseline), ncu launch list + full capture of K1 (C2)
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=10 > gpurun_out/pytest_gpu19.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu19.log
echo "== bench default (C3)"; timeout 600 python bench.py > gpurun_out/bench_C3_r01_v7.log 2>&1; echo "rc=$?"; tail -c 3000 gpurun_out/bench_C3_r01_v7.log
CMD="python bench.py --workload C2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
echo "== plain C2"; $CMD > gpurun_out/plain_C2_19.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_C2_19.log; exit 0; }
echo "== ncu launch list (C2, second solve)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_C2_r01.csv $CMD > gpurun_out/ncu_launches19.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_launches19.log | cut -c1-300; wc -l gpurun_out/launches_C2_r01.csv
echo "== ncu full on K1 (C2)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:syrk_dmma_kernel<0" -s 1 -c 2 -f -o gpurun_out/syrk_C2_r01 $CMD > gpurun_out/ncu_full19.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_full19.log | cut -c1-300; ls -la gpurun_out/*.ncu-rep
