#!/bin/bash
# Round 2, nineteenth GPU call (one GPU): probe workloads (module mixes of c3 and c4 swapped between their grids) for the A/B of the rolled Runge-Kutta stage loop in the kernels that also carry diffusion /
# sedimentation (default build) against the unrolled one (variants/unrolled), same box; parity first.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "probe workloads"
: > gpurun_out/sweep_roll2.jsonl
V=$PWD/mptrac_b200/_lib/variants
for wl in x3t x4s; do
  for v in unrolled default unrolled default; do
    if [ $v = default ]; then unset MPTRAC_B200_LIBDIR; else export MPTRAC_B200_LIBDIR=$V/$v; fi
    MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --workload $wl --no-cpu --no-exchange --steps 24 --warmup 3 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'workload':'$wl','variant':'$v','ms_per_step':d['ms_per_step'],'frac':d['roofline']['frac']}))" | tee -a gpurun_out/sweep_roll2.jsonl
  done
done
