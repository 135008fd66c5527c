#!/bin/bash
# Round 2, twentieth GPU call (one GPU): full ncu capture of the model-level advection kernel as it stands (record cache,
# fast weights), with the source view; bench of c2ml and of c3 on the final library.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
MPB_BENCH_NO_SUSTAIN=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:advect_levels -s 6 -c 2 -f -o gpurun_out/prof_c2ml \
  python bench.py --workload c2ml --steps 12 --warmup 3 --no-cpu --no-exchange > gpurun_out/ncu_c2ml.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_c2ml.ncu-rep
for wl in c2ml c3; do
  MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --workload $wl --no-cpu --no-exchange --steps 24 --warmup 3 2>/dev/null > gpurun_out/bench_${wl}_r02o.json
  python -c "import json; d=json.load(open('gpurun_out/bench_${wl}_r02o.json')); print('$wl', d['ms_per_step'], d['value'], d['roofline']['frac'])"
done
