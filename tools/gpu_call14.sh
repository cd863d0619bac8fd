#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/trace_solve.py C3 > gpurun_out/trace_C3_g.log 2>&1; echo "rc=$?"; head -2 gpurun_out/trace_C3_g.log; sed -n '/^14 gpu/,/^22 gpu/p' gpurun_out/trace_C3_g.log | cut -c1-260
timeout 300 python tools/trace_solve.py C2 > gpurun_out/trace_C2_g.log 2>&1; echo "rc=$?"; head -2 gpurun_out/trace_C2_g.log; tail -8 gpurun_out/trace_C2_g.log | cut -c1-260
