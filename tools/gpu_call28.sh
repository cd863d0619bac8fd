#!/bin/bash
# 1-GPU call: leaner pivot chain in potf2 / batched (tests, C2 / C3 / C4 numbers)
set -u
mkdir -p gpurun_out
echo "== pytest gpu (all, 1 GPU)"; timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=10 > gpurun_out/pytest_gpu28.log 2>&1; rc=$?; echo "rc=$rc"; tail -4 gpurun_out/pytest_gpu28.log
show() { grep -h '^{' $1 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['config'].get('iterations_per_solve'), d.get('phases_ms_per_solve'), d['roofline'].get('potrf_ms_per_launch'), d['roofline']['achieved'], d['e2e'])
" || tail -5 $1; }
echo "== bench C2"; python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C2_r01_v15.log 2>&1; echo "rc=$?"; show gpurun_out/bench_C2_r01_v15.log
echo "== bench C4"; python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C4_r01_v15.log 2>&1; echo "rc=$?"; show gpurun_out/bench_C4_r01_v15.log
echo "== bench C3"; timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_C3_r01_v15.log 2>&1; echo "rc=$?"; show gpurun_out/bench_C3_r01_v15.log
