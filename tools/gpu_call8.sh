#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== stress potrf m=16384 x40 (DMMA path only)"; timeout 300 python tools/stress_potrf.py 16384 40 00 > gpurun_out/stress40.log 2>&1; echo "rc=$?"; grep -c "maxdiff_vs_ref=0.000e+00" gpurun_out/stress40.log; tail -3 gpurun_out/stress40.log
for v in "0 0 1" "0 0 2" "1 0 1" "0 1 1"; do
echo "== in-solve verify C3 variant $v"; timeout 400 python tools/diag_potrf_insolve.py 16384 32768 $v > gpurun_out/insolve_$(echo $v | tr -d ' ').log 2>&1; echo "rc=$?"; tail -12 gpurun_out/insolve_$(echo $v | tr -d ' ').log | cut -c1-300
done
