#!/bin/bash
# Round 2, thirty-seventh GPU call (2 GPUs): smoke(), the multi-rank tests and the bench line at N=2 on the final library.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_dist.py tests/test_shim_trac.py -m gpu -q 2>&1 | tail -3
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 600 $L bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/bench_n2_r02t.json 2> gpurun_out/bench_n2_r02t.err; echo "rc=$?"
python -c "import json; d=json.load(open('gpurun_out/bench_n2_r02t.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k: (v.get('ms_per_step'), v.get('ms_transport_only'), v.get('error')) for k, v in d['exchange'].items()})"
timeout 300 $L bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
