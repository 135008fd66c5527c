#!/bin/bash
# Round 2, tenth GPU call (one GPU): whole gpu suite (one-pass output binning, uniform-time host stepping), the default bench
# line, sort cadence of the configs[2] / configs[3] workloads.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -8
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_c2.json')); print({k: d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['e2e']['d2h_bytes_per_step'], 'frac', d['roofline']['frac']); print({k: (v['ms_per_step'], v['ms_transport_only']) for k, v in d['exchange'].items()})"
: > gpurun_out/sweep_sort.jsonl
for wl in c3 c4; do for sd in 1200 1800 3600 7200; do
  MPB_BENCH_SORT_DT=$sd MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --workload $wl --no-cpu --steps 24 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'workload':'$wl','sort_dt':$sd,'ms_per_step':d['ms_per_step'],'b2b_ms':d['back_to_back']['ms_per_step'],'e2e_ms':d['e2e']['ms_per_step'],'frac':d['roofline']['frac']}))" | tee -a gpurun_out/sweep_sort.jsonl
done; done
