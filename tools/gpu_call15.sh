#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/diag_vectors.py 4096 8192 17 > gpurun_out/diagvec_C2.log 2>&1; echo "rc=$?"; cat gpurun_out/diagvec_C2.log | cut -c1-330
timeout 300 python tools/diag_vectors.py 2048 4096 15 > gpurun_out/diagvec_2k.log 2>&1; echo "rc=$?"; cat gpurun_out/diagvec_2k.log | cut -c1-330
