#!/bin/bash
# Round 2, second GPU call (one GPU): the whole gpu suite with the promoted + new tests (scale parity, team / peers on one device,
# shim variants), the diff_pbl diagnosis, the default bench line with its exchange sub-records.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rA --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "gpu suite rc=$?"
grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -30
timeout 300 python scripts/debug/diff_pbl_gpu.py > gpurun_out/diff_pbl_debug.txt 2>&1; tail -25 gpurun_out/diff_pbl_debug.txt
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; cat gpurun_out/bench_c2.json; tail -5 gpurun_out/bench_c2.err
