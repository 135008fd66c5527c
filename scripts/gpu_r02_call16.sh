#!/bin/bash
# Round 2, sixteenth GPU call (one GPU): A/B of the cheaper reciprocal in the metre -> degree divisor (production build), same box.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -q -x > gpurun_out/pytest_rcp.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/pytest_rcp.log
: > gpurun_out/sweep_rcp.jsonl
V=$PWD/mptrac_b200/_lib/variants
for wl in c2 c4; do
  for v in ieee_rcp default ieee_rcp default; do
    if [ $v = default ]; then unset MPTRAC_B200_LIBDIR; else export MPTRAC_B200_LIBDIR=$V/$v; fi
    MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --workload $wl --no-cpu --no-exchange --steps 48 --warmup 3 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'workload':'$wl','variant':'$v','ms_per_step':d['ms_per_step'],'frac':d['roofline']['frac']}))" | tee -a gpurun_out/sweep_rcp.jsonl
  done
done
