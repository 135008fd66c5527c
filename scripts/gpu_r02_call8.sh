#!/bin/bash
# Round 2, eighth GPU call (one GPU): the routed mixing exchange (team + two processes on one device), the scale tests, the
# default bench line.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_scale.py tests/test_gpu_parity.py tests/test_shim_trac.py -m gpu -q > gpurun_out/pytest_route.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/pytest_route.log
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print({k: d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'], d['roofline']['frac']); print(json.dumps(d['exchange'], indent=1)[:1500])"
