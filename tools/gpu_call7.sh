#!/bin/bash
set -u
mkdir -p gpurun_out
for m in 8192 16384; do
echo "== stress potrf m=$m"; timeout 400 python tools/stress_potrf.py $m 6 > gpurun_out/stress_$m.log 2>&1; echo "rc=$?"; cat gpurun_out/stress_$m.log | tail -32
done
