#!/bin/bash
# Round 2, final evidence call (one GPU): the whole gpu suite, the default bench line, its ncu launch list, one full ncu capture of
# the step kernel on C2 (the roofline's `traffic`), the reference arm.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; grep -E "passed|failed|FAILED|ERROR" gpurun_out/pytest_gpu.log | tail -5
timeout 600 python bench.py > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"; cat gpurun_out/bench_c2.json | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_c2.json 2> gpurun_out/bench_ref_c2.err; echo "reference arm rc=$?"; cut -c1-600 gpurun_out/bench_ref_c2.json
MPB_BENCH_NO_SUSTAIN=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv \
  python bench.py --steps 24 --warmup 3 --no-cpu --no-exchange > gpurun_out/launches_c2.log 2>&1; echo "launches rc=$?"
MPB_BENCH_NO_SUSTAIN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 14 -c 2 -f -o gpurun_out/prof_c2_final \
  python bench.py --steps 24 --warmup 3 --no-cpu --no-exchange > gpurun_out/ncu_c2.log 2>&1; echo "ncu rc=$?"
