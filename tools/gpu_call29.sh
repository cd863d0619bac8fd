#!/bin/bash
# 2-GPU call: slack-aware sweeps -- whole GPU suite incl. the world-2 sharded tests, C3 on 1 and 2 GPUs
set -u
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout 1200 python -m pytest tests -m gpu -q --tb=short --maxfail=10 > gpurun_out/pytest_gpu29.log 2>&1; rc=$?; echo "rc=$rc"; tail -5 gpurun_out/pytest_gpu29.log
show() { grep -h '^{' $1 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config'].get('iterations_per_solve'), d.get('phases_ms_per_solve'), d['roofline'].get('potrf_ms_per_launch'), d['roofline']['achieved'])
" || tail -5 $1; }
echo "== bench C3 1 GPU"; timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_C3_r01_v16.log 2>&1; echo "rc=$?"; show gpurun_out/bench_C3_r01_v16.log
echo "== bench C3 x2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e > gpurun_out/bench_C3_n2_r01_v16.log 2>&1; echo "rc=$?"; show gpurun_out/bench_C3_n2_r01_v16.log
