#!/bin/bash
# Round 2, thirteenth GPU call (EIGHT GPUs, short): bench line at N=8 with the final routed exchange, N=4 for the scaling curve.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus $n --steps 24 --warmup 3 > gpurun_out/bench_n${n}_final.json 2> gpurun_out/bench_n${n}_final.err; echo "bench n$n rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n${n}_final.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step']); print({k: (v.get('ms_per_step'), v.get('ms_transport_only'), v.get('error')) for k, v in d['exchange'].items()})" || tail -20 gpurun_out/bench_n${n}_final.err
done
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q 2>&1 | tail -2
