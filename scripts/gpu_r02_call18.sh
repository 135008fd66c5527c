#!/bin/bash
# Round 2, eighteenth GPU call (one GPU): A/B of the rolled Runge-Kutta stage loop in the kernels that also carry diffusion /
# sedimentation (default build) against the unrolled one (variants/unrolled), same box; parity first.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_modules.py -m gpu -q -x > gpurun_out/pytest_roll.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/pytest_roll.log
: > gpurun_out/sweep_roll.jsonl
V=$PWD/mptrac_b200/_lib/variants
for wl in c4 c3 c2; do
  for v in unrolled default unrolled default; do
    if [ $v = default ]; then unset MPTRAC_B200_LIBDIR; else export MPTRAC_B200_LIBDIR=$V/$v; fi
    MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --workload $wl --no-cpu --no-exchange --steps 24 --warmup 3 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'workload':'$wl','variant':'$v','ms_per_step':d['ms_per_step'],'frac':d['roofline']['frac']}))" | tee -a gpurun_out/sweep_roll.jsonl
  done
done
