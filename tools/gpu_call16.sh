#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python tools/accuracy_probe.py 2048 8192 > gpurun_out/accuracy.log 2>&1; echo "rc=$?"; cat gpurun_out/accuracy.log
