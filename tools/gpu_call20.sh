#!/bin/bash
# 1-GPU call: slack-structured SYRK (tests + C3/C2 numbers), ncu --set full of K1 / potf2_inv / trsm_blocked / update
set -u
mkdir -p gpurun_out
echo "== pytest gpu (kernels, solve, batched, cpp)"; timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_solve.py tests/test_gpu_batched.py tests/test_cpp_host.py -m gpu -q --tb=short --maxfail=10 > gpurun_out/pytest_gpu20.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/pytest_gpu20.log
echo "== bench C3"; timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/bench_C3_r01_v8.log 2>&1; echo "rc=$?"; tail -c 2600 gpurun_out/bench_C3_r01_v8.log
CMD="python bench.py --workload C2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
echo "== plain C2"; $CMD > gpurun_out/bench_C2_r01_v8.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/bench_C2_r01_v8.log; exit 0; }
tail -c 1500 gpurun_out/bench_C2_r01_v8.log
echo "== ncu full K1 (C2)"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:syrk_dmma_kernel<\(int\)0' -s 1 -c 1 -f -o gpurun_out/syrk_C2_r01 $CMD > gpurun_out/ncu_full20a.log 2>&1
echo "rc=$?"; grep -E "PROF|WARN|ERR" gpurun_out/ncu_full20a.log | tail -3 | cut -c1-200
echo "== ncu full potf2_inv + trsm_blocked (C2)"
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:potf2_inv_kernel|trsm_blocked_kernel' -s 10 -c 4 -f -o gpurun_out/panel_C2_r01 $CMD > gpurun_out/ncu_full20b.log 2>&1
echo "rc=$?"; grep -E "PROF|WARN|ERR" gpurun_out/ncu_full20b.log | tail -3 | cut -c1-200
echo "== ncu full trailing update (C2)"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:syrk_dmma_kernel<\(int\)1' -s 4 -c 1 -f -o gpurun_out/update_C2_r01 $CMD > gpurun_out/ncu_full20c.log 2>&1
echo "rc=$?"; grep -E "PROF|WARN|ERR" gpurun_out/ncu_full20c.log | tail -3 | cut -c1-200
ls -la gpurun_out/*.ncu-rep
