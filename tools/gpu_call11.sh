#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/trace_solve.py C3 > gpurun_out/trace_C3_a.log 2>&1; echo "rc=$?"; cat gpurun_out/trace_C3_a.log | cut -c1-220
timeout 300 python tools/trace_solve.py C3 solve_impl=1 > gpurun_out/trace_C3_b.log 2>&1; echo "rc=$?"; head -3 gpurun_out/trace_C3_b.log; tail -12 gpurun_out/trace_C3_b.log | cut -c1-220
timeout 300 python tools/trace_solve.py C3 trsm_impl=1 > gpurun_out/trace_C3_c.log 2>&1; echo "rc=$?"; head -3 gpurun_out/trace_C3_c.log; tail -8 gpurun_out/trace_C3_c.log | cut -c1-220
