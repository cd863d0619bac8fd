#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python tools/stress_potrf.py 16384 120 0,0,0 0,0,1 0,2,0 0,3,0 1,0,0 > gpurun_out/stress9.log 2>&1; echo "rc=$?"; cut -c1-400 gpurun_out/stress9.log | tail -40
