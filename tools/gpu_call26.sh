#!/bin/bash
# 1-GPU call: batched e2e host-time diagnostic, batched tests, ncu --set full of K1 at the C3 shape (traffic),
# launch list of the default bench command (C3)
set -u
mkdir -p gpurun_out
echo "== batched e2e timing"; timeout 300 python tools/time_batched_e2e.py > gpurun_out/time_batched26.log 2>&1; echo "rc=$?"; cat gpurun_out/time_batched26.log | tail -12
echo "== pytest batched + abi"; timeout 600 python -m pytest tests/test_gpu_batched.py tests/test_abi.py -q --tb=short > gpurun_out/pytest_gpu26.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu26.log
CMD="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
echo "== plain C3"; $CMD > gpurun_out/plain_C3_26.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_C3_26.log; exit 0; }
tail -c 600 gpurun_out/plain_C3_26.log
echo "== ncu full K1 (C3)"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:syrk_dmma_kernel<\(int\)0' -s 1 -c 1 -f -o gpurun_out/syrk_C3_r01 $CMD > gpurun_out/ncu_full26.log 2>&1
echo "rc=$?"; grep -E "PROF|WARN|ERR" gpurun_out/ncu_full26.log | tail -2 | cut -c1-200
echo "== ncu launch list (C3, default command, one solve)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_C3_r01_v13.csv $CMD > gpurun_out/ncu_launches26.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches_C3_r01_v13.csv
