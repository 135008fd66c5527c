#!/bin/bash
# Round 2, fortieth GPU call (one GPU): the own parcel order with the streaming key pass -- its test, the whole GPU suite,
# and the workloads with it (default) and without (MPTRAC_B200_PRIVATE_ORDER=0), same box.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "own_parcel_order" 2>&1 | tail -12
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r02w.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/pytest_gpu_r02w.log
: > gpurun_out/sweep_own_order2.jsonl
for wl in c3 c4; do
  for v in 0 1 0 1; do
    MPTRAC_B200_PRIVATE_ORDER=$v MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --workload $wl --no-cpu --no-exchange --steps 24 --warmup 3 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'workload':'$wl','own_order':$v,'ms_per_step':d['ms_per_step'],'frac':d['roofline']['frac']}))" | tee -a gpurun_out/sweep_own_order2.jsonl
  done
done
for v in 0 1; do
  MPTRAC_B200_PRIVATE_ORDER=$v MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --no-cpu --steps 12 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'own_order':$v, **{k: [round(x['ms_per_step'],4), round(x['ms_transport_only'],4)] for k, x in d['exchange'].items()}}))" | tee -a gpurun_out/sweep_own_order2.jsonl
done
