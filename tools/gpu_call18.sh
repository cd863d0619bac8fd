#!/bin/bash
# 2-GPU call: whole GPU test-suite, sharded C3 / C2 benches
set -u
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=10 > gpurun_out/pytest_gpu18.log 2>&1; rc=$?; echo "rc=$rc"; tail -8 gpurun_out/pytest_gpu18.log
echo "== bench C3 sharded x2"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --workload C3 --steps 2 --warmup 1 > gpurun_out/bench_C3_n2.log 2>&1; echo "rc=$?"; tail -c 2300 gpurun_out/bench_C3_n2.log
echo "== bench C2 sharded x2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --workload C2 --steps 3 --warmup 3 > gpurun_out/bench_C2_n2.log 2>&1; echo "rc=$?"; tail -c 2300 gpurun_out/bench_C2_n2.log
