#!/bin/bash
# 2-GPU call: full GPU test-suite (incl. the world-2 sharded tests), sharded C3 bench, batched C4 bench
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpu_info2.txt 2>&1
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short --maxfail=8 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench C4 (1 GPU)"; timeout 600 python bench.py --workload C4 --steps 3 --warmup 3 > gpurun_out/bench_C4_n1.log 2>&1; echo "rc=$?"; tail -c 2000 gpurun_out/bench_C4_n1.log
echo "== bench C2 sharded x2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload C2 --steps 3 --warmup 3 > gpurun_out/bench_C2_n2.log 2>&1; echo "rc=$?"; tail -c 2500 gpurun_out/bench_C2_n2.log
echo "== bench C3 sharded x2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload C3 --steps 1 --warmup 1 --no-e2e > gpurun_out/bench_C3_n2.log 2>&1; echo "rc=$?"; tail -c 2500 gpurun_out/bench_C3_n2.log
echo "== bench C4 x2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload C4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C4_n2.log 2>&1; echo "rc=$?"; tail -c 2000 gpurun_out/bench_C4_n2.log
