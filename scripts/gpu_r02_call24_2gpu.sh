#!/bin/bash
# Round 2, twenty-fourth GPU call (2 GPUs): multi-rank tests and the (untraced, then traced) exchange records with the
# answers of the routed exchange sent as box means.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_scale.py -m gpu -q -x 2>&1 | tail -3
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
MPB_BENCH_NO_SUSTAIN=1 timeout 600 $L bench.py --gpus 2 --steps 12 --warmup 3 --no-cpu > gpurun_out/bench_n2_means.json 2> gpurun_out/bench_n2_means.err; echo "rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench_n2_means.json')); print({k: (v.get('ms_per_step'), v.get('ms_transport_only'), v.get('error')) for k, v in d['exchange'].items()})"
MPTRAC_B200_TRACE_MIXING=1 MPB_BENCH_NO_SUSTAIN=1 timeout 600 $L bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu > gpurun_out/trace_mix_2gpu.json 2> gpurun_out/trace_mix_2gpu.err
grep "mixing trace" gpurun_out/trace_mix_2gpu.err
