#!/bin/bash
# Round 2, sixth GPU call (one GPU): the whole gpu suite after the restatement of diff_turb / diff_pbl, the padded box records
# and the variant-test fixes; the unmodified trac drop-in at C2 scale, warm, with / without the context warm-up and with the
# strict library.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "gpu suite rc=$?"
grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -12
timeout 1200 scripts/trac_dropin_bench.sh > gpurun_out/trac_dropin.txt 2>&1; echo "dropin rc=$?"; cat gpurun_out/trac_dropin.txt | tail -60
