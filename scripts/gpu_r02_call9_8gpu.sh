#!/bin/bash
# Round 2, ninth GPU call (EIGHT GPUs, short): the routed mixing exchange on distinct devices -- tests, then the bench line at N=8.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -rA > gpurun_out/pytest_8gpu.log 2>&1; echo "multi-rank tests rc=$?"; grep -E "passed|failed|PASSED|FAILED|SKIPPED" gpurun_out/pytest_8gpu.log | tail -8
for n in 8 2; do
MPB_BENCH_EXCHANGE=peers timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus $n --steps 24 --warmup 3 > gpurun_out/bench_n${n}_routed.json 2> gpurun_out/bench_n${n}_routed.err; echo "bench n$n rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n${n}_routed.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step']); print(json.dumps(d.get('exchange'), indent=1))" || tail -20 gpurun_out/bench_n${n}_routed.err
done
