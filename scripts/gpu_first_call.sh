#!/bin/bash
# First GPU call of a round (under gpurun, one GPU, ~4 min of box time): the verified suite, the parked tests of kernels
# that have not run on a device yet, the headline bench line and its launch list.  Everything lands in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- scripts/gpu_first_call.sh
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/pytest_gpu.log
# no -x here: every parked test reports on its own
timeout 900 python -m pytest tests -m gpu_pending -q -rA > gpurun_out/pytest_gpu_pending.log 2>&1; echo "gpu_pending rc=$?"; tail -30 gpurun_out/pytest_gpu_pending.log
scripts/gpu_check.sh bench launches
