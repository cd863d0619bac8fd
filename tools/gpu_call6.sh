#!/bin/bash
# 2-GPU call: new potf2_inv / pipelined solve kernels (tests), rank-divergence diagnostic, quick C2/C3 numbers
set -u
mkdir -p gpurun_out
echo "== pytest kernels+solve"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py tests/test_gpu_batched.py -m gpu -q --tb=short --maxfail=10 -x > gpurun_out/pytest_gpu6.log 2>&1; rc=$?; echo "rc=$rc"; tail -30 gpurun_out/pytest_gpu6.log
if [ $rc -ne 0 ]; then echo "tests failed; skipping the rest"; exit 0; fi
echo "== bench C2 1 GPU"; timeout 300 python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_C2_v3.log 2>&1; echo "rc=$?"; tail -c 1800 gpurun_out/bench_C2_v3.log
echo "== diag 8192x16384"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 tools/diag_sharded.py 8192 16384 > gpurun_out/diag_8k.log 2>&1; echo "rc=$?"; grep -v "^[0-9]* [0-9]" gpurun_out/diag_8k.log | tail -25
echo "== diag C3"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 tools/diag_sharded.py 16384 32768 > gpurun_out/diag_C3.log 2>&1; echo "rc=$?"; grep -v "^[0-9]* [0-9]" gpurun_out/diag_C3.log | tail -25
echo "== bench C3 1 GPU"; timeout 400 python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_C3_v3.log 2>&1; echo "rc=$?"; tail -c 1800 gpurun_out/bench_C3_v3.log
