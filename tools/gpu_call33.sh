#!/bin/bash
# 1-GPU call: compute-sanitizer memcheck over small end-to-end invocations of every single-GPU kernel family
set -u
mkdir -p gpurun_out
python tools/sanitize_smoke.py > gpurun_out/sanitize_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain.log; exit 0; }
tail -3 gpurun_out/sanitize_plain.log
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_smoke.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/sanitize_memcheck.log | cut -c1-220
