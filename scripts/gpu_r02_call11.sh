#!/bin/bash
# Round 2, eleventh GPU call (one GPU): how sensitive is the step kernel to occupancy?  Unused dynamic shared memory lowers the
# resident blocks per SM from 4 to 3, 2 and 1 (16 -> 12 -> 8 -> 4 warps per SM).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k host_resident > gpurun_out/pytest_host.log 2>&1; echo "host-resident tests rc=$?"; tail -2 gpurun_out/pytest_host.log
: > gpurun_out/sweep_occupancy.jsonl
for wl in c2 c4; do for pad in 0 60000 90000 200000; do
  MPTRAC_B200_PAD_SMEM=$pad MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --workload $wl --no-cpu --no-exchange --steps 24 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'workload':'$wl','pad_smem':$pad,'ms_per_step':d['ms_per_step'],'frac':d['roofline']['frac']}))" | tee -a gpurun_out/sweep_occupancy.jsonl
done; done
