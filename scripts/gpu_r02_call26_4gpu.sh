#!/bin/bash
# Round 2, twenty-sixth GPU call (4 GPUs, short): the bench line at N=4 on the final library, for the exchange table.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $L bench.py --gpus 4 --steps 24 --warmup 3 --no-cpu > gpurun_out/bench_n4_r02q.json 2> gpurun_out/bench_n4_r02q.err; echo "bench n4 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n4_r02q.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step']); print({k: (v.get('ms_per_step'), v.get('ms_transport_only'), v.get('error')) for k, v in d['exchange'].items()})" || tail -20 gpurun_out/bench_n4_r02q.err
