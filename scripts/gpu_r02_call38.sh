#!/bin/bash
# Round 2, thirty-eighth GPU call (one GPU): module_timesteps folded into the sort's key pass -- the whole GPU suite, then the
# default bench line (exchange record c5 sorts every step).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02u.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_gpu_r02u.log
timeout 600 python bench.py --no-cpu > gpurun_out/bench_c2_r02u.json 2> gpurun_out/bench_c2_r02u.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_c2_r02u.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'])
for k,v in d['exchange'].items(): print(k, v['ms_per_step'], v['ms_transport_only'])"
