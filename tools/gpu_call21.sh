#!/bin/bash
# 2-GPU call: distributed (panel-broadcast) Cholesky -- sharded tests, C3 / C2 sharded benches with and without it
set -u
mkdir -p gpurun_out
echo "== pytest sharded"; timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --tb=short > gpurun_out/pytest_gpu21.log 2>&1; rc=$?; echo "rc=$rc"; tail -15 gpurun_out/pytest_gpu21.log
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:3}" > gpurun_out/$2 2>&1; echo "rc=$?"; grep -h '^{' gpurun_out/$2 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['value'], d['ms_per_step'], d['config']['iterations_per_solve'], d['phases_ms_per_solve'], d['roofline']['potrf_ms_per_launch'])
" || tail -5 gpurun_out/$2; }
echo "== bench C3 x2 (dist potrf)"; run 29541 bench_C3_n2_r01_v9.log --workload C3 --steps 2 --warmup 1 --no-e2e
echo "== bench C2 x2 (dist potrf)"; run 29542 bench_C2_n2_r01_v9.log --workload C2 --steps 3 --warmup 3 --no-e2e
echo "== bench C3 x2 (replicated potrf)"; run 29543 bench_C3_n2_r01_v9_repl.log --workload C3 --steps 1 --warmup 1 --no-e2e --potrf-dist 0
