#!/bin/bash
# 8-GPU call: C3 and C5 with the final kernels (slack-aware sweeps, packed all-reduce, lean panel kernels)
set -u
mkdir -p gpurun_out
show() { grep -h '^{' $1 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], d['config']['iterations_per_solve'], d['phases_ms_per_solve'], d['roofline']['potrf_ms_per_launch'], d['roofline']['achieved'], d['e2e'])
" || tail -5 $1; }
echo "== bench C3 x8"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --workload C3 --steps 3 --warmup 3 > gpurun_out/bench_C3_n8_r01_v17.log 2>&1; echo "rc=$?"; show gpurun_out/bench_C3_n8_r01_v17.log; tail -2 gpurun_out/bench_C3_n8_r01_v17.log | cut -c1-200
echo "== bench C5 x8"; timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29582 bench.py --gpus 8 --workload C5 --steps 1 --warmup 1 --no-e2e > gpurun_out/bench_C5_n8_r01_v17.log 2>&1; echo "rc=$?"; show gpurun_out/bench_C5_n8_r01_v17.log; tail -2 gpurun_out/bench_C5_n8_r01_v17.log | cut -c1-200
