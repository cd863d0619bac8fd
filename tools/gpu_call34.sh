#!/bin/bash
# 1-GPU call (last of the round): look-ahead Cholesky -- kernel + solve tests, smoke(), C2 / C3 numbers
set -u
mkdir -p gpurun_out
echo "== pytest gpu (kernels, solve, cpp)"; timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py tests/test_cpp_host.py -m gpu -q --tb=short --maxfail=10 > gpurun_out/pytest_gpu34.log 2>&1; rc=$?; echo "rc=$rc"; tail -4 gpurun_out/pytest_gpu34.log
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
show() { grep -h '^{' $1 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['config'].get('iterations_per_solve'), d.get('phases_ms_per_solve'), d['roofline'].get('potrf_ms_per_launch'), d['roofline']['achieved'])
" || tail -5 $1; }
echo "== bench C2"; python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_C2_r01_v18.log 2>&1; echo "rc=$?"; show gpurun_out/bench_C2_r01_v18.log
echo "== bench C3"; timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_C3_r01_v18.log 2>&1; echo "rc=$?"; show gpurun_out/bench_C3_r01_v18.log
