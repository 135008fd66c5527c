#!/bin/bash
# Round 2, fourth GPU call (TWO GPUs): the multi-rank tests on distinct devices (team, peers over CUDA IPC, NCCL), the shim's team
# of devices, the quad variant's tests, the bench under torchrun with both transports of the exchange step, the record-cache
# flavour of the model-level advection.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_2gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_quad.py "tests/test_shim_trac.py::test_trac_dt_test_on_a_team_of_devices_is_bit_identical" -m gpu -q -rA > gpurun_out/pytest_2gpu.log 2>&1; echo "2-GPU tests rc=$?"
grep -E "passed|failed|PASSED|FAILED|SKIPPED|Error" gpurun_out/pytest_2gpu.log | tail -40
for tr in peers nccl; do
  MPB_BENCH_EXCHANGE=$tr timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 24 --warmup 3 > gpurun_out/bench_n2_$tr.json 2> gpurun_out/bench_n2_$tr.err; echo "bench n2 $tr rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_n2_$tr.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step']); print(json.dumps(d.get('exchange'), indent=1))" || tail -20 gpurun_out/bench_n2_$tr.err
done
LC=$PWD/mptrac_b200/_lib/variants/f128m4_lc
for v in default lc; do
  if [ $v = lc ]; then export MPTRAC_B200_LIBDIR=$LC; fi
  MPB_BENCH_NO_SUSTAIN=1 timeout 300 python bench.py --workload c2ml --no-cpu --steps 36 --warmup 3 > gpurun_out/bench_c2ml_$v.json 2> gpurun_out/bench_c2ml_$v.err
  python -c "import json; d=json.load(open('gpurun_out/bench_c2ml_$v.json')); print('c2ml $v: ms/step', d['ms_per_step'], 'value', d['value'])"
done
