#!/bin/bash
# Round 2, forty-second GPU call (2 GPUs, short): the bench line at N=2 on the final library (the configs[3] share runs with the
# engine's own parcel order on every rank).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
MPB_BENCH_NO_SUSTAIN=1 timeout 200 $L bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu > gpurun_out/bench_n2_r02v.json 2> gpurun_out/bench_n2_r02v.err; echo "rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench_n2_r02v.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k: (v.get('ms_per_step'), v.get('ms_transport_only'), v.get('error')) for k, v in d['exchange'].items()})"
