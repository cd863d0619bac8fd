#!/bin/bash
# 1-GPU call: new TRSM (DMMA) + potf2 (fused factor/inverse) kernels: tests, C2 / C3 numbers, launch list
set -u
mkdir -p gpurun_out
echo "== pytest gpu (kernels, solve)"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py tests/test_cpp_host.py -m gpu -q --tb=short --maxfail=10 > gpurun_out/pytest_gpu22.log 2>&1; rc=$?; echo "rc=$rc"; tail -12 gpurun_out/pytest_gpu22.log
show() { grep -h '^{' $1 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['config']['iterations_per_solve'], d['phases_ms_per_solve'], d['roofline']['potrf_ms_per_launch'], d['roofline']['achieved'])
" || tail -5 $1; }
CMD="python bench.py --workload C2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
echo "== bench C2"; $CMD > gpurun_out/bench_C2_r01_v10.log 2>&1; echo "rc=$?"; show gpurun_out/bench_C2_r01_v10.log
echo "== bench C3"; timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_C3_r01_v10.log 2>&1; echo "rc=$?"; show gpurun_out/bench_C3_r01_v10.log
echo "== ncu launch list (C2)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches_C2_r01_v10.csv $CMD > gpurun_out/ncu_launches22.log 2>&1
echo "rc=$?"; wc -l gpurun_out/launches_C2_r01_v10.csv
