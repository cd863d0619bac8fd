#!/bin/bash
# 8-GPU call: C3 strong scaling point and C5 (device-generated, column-sharded, distributed Cholesky)
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/gpu_info8.txt 2>&1
show() { grep -h '^{' $1 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['iterations_per_solve'], d['phases_ms_per_solve'], d['roofline']['potrf_ms_per_launch'], d['roofline']['achieved'])
" || tail -5 $1; }
echo "== bench C3 x8"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --workload C3 --steps 2 --warmup 1 --no-e2e > gpurun_out/bench_C3_n8_r01_v11.log 2>&1; echo "rc=$?"; show gpurun_out/bench_C3_n8_r01_v11.log; tail -3 gpurun_out/bench_C3_n8_r01_v11.log | cut -c1-300
echo "== bench C5 x8"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 8 --workload C5 --steps 1 --warmup 1 --no-e2e > gpurun_out/bench_C5_n8_r01_v11.log 2>&1; echo "rc=$?"; show gpurun_out/bench_C5_n8_r01_v11.log; tail -3 gpurun_out/bench_C5_n8_r01_v11.log | cut -c1-300
