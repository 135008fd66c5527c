#!/bin/bash
# Round 2, fourteenth GPU call (one GPU): look-ahead prefetch of the destination cell -- off / into L1 / into L2 -- on C2, C3, C4.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/pytest_pf.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/pytest_pf.log
: > gpurun_out/sweep_prefetch.jsonl
V=$PWD/mptrac_b200/_lib/variants
for wl in c2 c4 c3; do
  for v in pf0 default pfL2 pf0 default; do
    if [ $v = default ]; then unset MPTRAC_B200_LIBDIR; else export MPTRAC_B200_LIBDIR=$V/$v; fi
    MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --workload $wl --no-cpu --no-exchange --steps 36 --warmup 3 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'workload':'$wl','variant':'$v','ms_per_step':d['ms_per_step'],'b2b_ms':d['back_to_back']['ms_per_step'],'frac':d['roofline']['frac']}))" | tee -a gpurun_out/sweep_prefetch.jsonl
  done
done
