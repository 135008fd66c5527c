#!/bin/bash
# Round 2, twenty-eighth GPU call (one GPU): full ncu capture of the binning kernel of configs[3]'s gridded output.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
MPB_BENCH_NO_SUSTAIN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:grid_bin -s 4 -c 2 -f -o gpurun_out/prof_grid_bin \
  python bench.py --workload c4g --steps 6 --warmup 3 --no-cpu > gpurun_out/ncu_grid_bin.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_grid_bin.ncu-rep
