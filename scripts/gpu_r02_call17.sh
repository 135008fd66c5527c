#!/bin/bash
# Round 2, seventeenth GPU call (one GPU): multi-rank tests with the all-reduce transport over gloo on one device; launch list of
# the configs[4] step (sort + mixing kernels).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -rA > gpurun_out/pytest_dist.log 2>&1; echo "tests rc=$?"; grep -E "PASSED|FAILED|SKIPPED|passed|failed|Error" gpurun_out/pytest_dist.log | tail -10
MPB_BENCH_NO_SUSTAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv \
  python bench.py --workload c5 --steps 6 --warmup 3 --no-cpu > gpurun_out/launches_c5.log 2>&1; echo "launches c5 rc=$?"
MPB_BENCH_NO_SUSTAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c4g.csv \
  python bench.py --workload c4g --steps 6 --warmup 3 --no-cpu > gpurun_out/launches_c4g.log 2>&1; echo "launches c4g rc=$?"
