#!/bin/bash
# Round 2, thirty-fifth GPU call (one GPU): the whole GPU suite and the bench lines of every workload on the final library.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r02t.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_gpu_r02t.log
timeout 600 python bench.py > gpurun_out/bench_c2_r02t.json 2> gpurun_out/bench_c2_r02t.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_c2_r02t.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e'].get('ms_per_step'), d['roofline']['frac'], d['clocks'], d['cpu_baseline']['value'])
for k,v in d['exchange'].items(): print(k, v['ms_per_step'], v['ms_transport_only'])"
for wl in c3 c4 c2ml; do
  timeout 400 python bench.py --workload $wl --no-cpu --no-exchange 2>/dev/null > gpurun_out/bench_${wl}_r02t.json
  python -c "import json; d=json.load(open('gpurun_out/bench_${wl}_r02t.json')); print('$wl', d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'])"
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r02t.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/bench_ref_r02t.json')); print('reference', d['value'], d['cpu_baseline'])"
