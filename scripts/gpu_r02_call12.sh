#!/bin/bash
# Round 2, twelfth GPU call (one GPU): routed exchange after the route / serve changes, model-level advection with reciprocal
# weights and the hoisted metre -> degree divisor.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_scale.py tests/test_gpu_parity.py tests/test_shim_trac.py -m gpu -q > gpurun_out/pytest_route.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_route.log
MPB_BENCH_NO_SUSTAIN=1 timeout 300 python bench.py --workload c2ml --no-cpu --steps 36 --warmup 3 > gpurun_out/bench_c2ml.json 2> gpurun_out/bench_c2ml.err
python -c "import json; d=json.load(open('gpurun_out/bench_c2ml.json')); print('c2ml: ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['ms_per_step'])"
