#!/bin/bash
# Round 2, thirty-fourth GPU call (one GPU): how much does the reference's sort key cost on a 0..360 met grid?  It clamps the
# longitudes of the western half of the parcels into column 0, so those parcels are sorted by latitude and level only.
# Probe: the same workloads on a grid labelled -180..180 (every parcel sorted by its full cell).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/sweep_lon_axis.jsonl
for wl in c2 c3 c4; do
  for ax in 0 -180 0 -180; do
    MPB_BENCH_LON_AXIS=$ax MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --workload $wl --no-cpu --no-exchange --steps 24 --warmup 3 2>/dev/null \
      | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'workload':'$wl','lon_axis':'$ax','ms_per_step':d['ms_per_step'],'frac':d['roofline']['frac']}))" | tee -a gpurun_out/sweep_lon_axis.jsonl
  done
done
