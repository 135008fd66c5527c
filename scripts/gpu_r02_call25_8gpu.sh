#!/bin/bash
# Round 2, twenty-fifth GPU call (EIGHT GPUs, short): the bench line at N=8 on the final library (two-level slot allocation,
# answers as box means), then the phase times of module_mixing on 8 ranks.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $L bench.py --gpus 8 --steps 24 --warmup 3 > gpurun_out/bench_n8_r02q.json 2> gpurun_out/bench_n8_r02q.err; echo "bench n8 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n8_r02q.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step']); print({k: (v.get('ms_per_step'), v.get('ms_transport_only'), v.get('error')) for k, v in d['exchange'].items()})" || tail -20 gpurun_out/bench_n8_r02q.err
MPTRAC_B200_TRACE_MIXING=1 MPB_BENCH_NO_SUSTAIN=1 timeout 300 $L bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu > gpurun_out/trace_mix_8gpu.json 2> gpurun_out/trace_mix_8gpu.err
grep "mixing trace" gpurun_out/trace_mix_8gpu.err
