#!/bin/bash
# Round 2, thirty-first GPU call (one GPU): launch lists of the configs[3] step INCLUDING cell sorts (27 steps: sorts at steps
# 12 and 24), previous library (variants/old_bin) and current one: the binning kernel before and after a sort.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
V=$PWD/mptrac_b200/_lib/variants
for v in old_bin default; do
  if [ $v = default ]; then unset MPTRAC_B200_LIBDIR; else export MPTRAC_B200_LIBDIR=$V/$v; fi
  MPB_BENCH_NO_SUSTAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4g_sorted_$v.csv \
    python bench.py --workload c4g --steps 24 --warmup 3 --no-cpu --no-exchange > gpurun_out/launches_c4g_$v.log 2>&1; echo "launches $v rc=$?"
done
