#!/bin/bash
# 1-GPU call: batched kernel rewrite (tests + C4 bench), whole GPU suite, ncu of the new potf2_inv, C5 on one GPU
set -u
mkdir -p gpurun_out
echo "== pytest gpu (all, 1 GPU)"; timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=10 > gpurun_out/pytest_gpu24.log 2>&1; rc=$?; echo "rc=$rc"; tail -8 gpurun_out/pytest_gpu24.log
echo "== bench C4"; timeout 600 python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_C4_r01_v12.log 2>&1; echo "rc=$?"; tail -c 1600 gpurun_out/bench_C4_r01_v12.log
CMD="python bench.py --workload C2 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e"
echo "== plain C2"; $CMD > gpurun_out/plain_C2_24.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_C2_24.log; }
echo "== ncu full potf2_inv (C2)"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:potf2_inv_kernel' -s 5 -c 2 -f -o gpurun_out/potf2_C2_r01_v12 $CMD > gpurun_out/ncu_full24.log 2>&1
echo "rc=$?"; grep -E "PROF|WARN|ERR" gpurun_out/ncu_full24.log | tail -2 | cut -c1-200
echo "== bench C5 on 1 GPU"; timeout 900 python bench.py --workload C5 --steps 1 --warmup 0 --no-e2e > gpurun_out/bench_C5_n1_r01_v12.log 2>&1; echo "rc=$?"; tail -c 2500 gpurun_out/bench_C5_n1_r01_v12.log
