#!/bin/bash
# Round 2, last GPU call (one GPU): the whole GPU suite and smoke() on the final library.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 170 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_gpu_final.log
timeout 30 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
