#!/bin/bash
# Round 2, seventh GPU call (EIGHT GPUs, short): the multi-rank tests on distinct devices, the bench line at N=8 with the exchange
# sub-records (peer-memory transport, then NCCL), the unmodified trac on a team of 8 devices.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_8gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -rA > gpurun_out/pytest_8gpu.log 2>&1; echo "multi-rank tests rc=$?"; grep -E "passed|failed|PASSED|FAILED|SKIPPED" gpurun_out/pytest_8gpu.log | tail -8
for tr in peers nccl; do
  MPB_BENCH_EXCHANGE=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus 8 --steps 24 --warmup 3 > gpurun_out/bench_n8_$tr.json 2> gpurun_out/bench_n8_$tr.err; echo "bench n8 $tr rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_n8_$tr.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e'].get('host_link')); print(json.dumps(d.get('exchange'), indent=1))" || tail -20 gpurun_out/bench_n8_$tr.err
done
timeout 600 scripts/trac_dropin_bench.sh > gpurun_out/trac_dropin_8gpu.txt 2>&1; echo "dropin rc=$?"; grep -A8 "== gpu_team\|== gpu:" gpurun_out/trac_dropin_8gpu.txt | head -40; tail -4 gpurun_out/trac_dropin_8gpu.txt
