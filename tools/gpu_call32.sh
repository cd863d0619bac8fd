#!/bin/bash
# 1-GPU call: the default bench command, final build of the round (e2e + cpu_baseline + roofline.traffic)
set -u
mkdir -p gpurun_out
timeout 700 python bench.py > gpurun_out/bench_C3_r01_final.log 2>&1; echo "rc=$?"; tail -c 3500 gpurun_out/bench_C3_r01_final.log
