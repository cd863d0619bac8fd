#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/trace_solve.py C3 solve_impl=2 > gpurun_out/trace_C3_d.log 2>&1; echo "rc=$?"; head -3 gpurun_out/trace_C3_d.log; tail -9 gpurun_out/trace_C3_d.log | cut -c1-220
timeout 300 python tools/trace_solve.py C3 solve_impl=2 trsm_impl=1 > gpurun_out/trace_C3_e.log 2>&1; echo "rc=$?"; head -3 gpurun_out/trace_C3_e.log; tail -9 gpurun_out/trace_C3_e.log | cut -c1-220
