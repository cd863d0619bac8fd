#!/bin/bash
# 2-GPU call: the whole GPU suite on the final tree of the round
set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --tb=short --maxfail=10 > gpurun_out/pytest_gpu35.log 2>&1; rc=$?; echo "rc=$rc"; tail -5 gpurun_out/pytest_gpu35.log
