#!/bin/bash
# Round 2, thirty-sixth GPU call (one GPU): the binning test at the cell faces, then the rest of the GPU suite.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cell_faces" 2>&1 | tail -15
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02t.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/pytest_gpu_r02t.log
