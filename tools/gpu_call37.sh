#!/bin/bash
# last call of the round: smoke() + the look-ahead / structure tests on the final tree
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 60 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "lookahead or syrk_adat" > gpurun_out/pytest_gpu37.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/pytest_gpu37.log
