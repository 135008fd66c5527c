#!/bin/bash
# Round 2, thirtieth GPU call (one GPU): binning kernel with a shared-memory table of partial sums per block, lane-private partial sums, the single-precision level decision and the pinned landing area of the fetch
# (default) against the previous library (variants/old_bin), same box; parity first; launch list of the configs[3] step.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_modules.py tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/pytest_bin.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/pytest_bin.log
: > gpurun_out/sweep_bin3.jsonl
V=$PWD/mptrac_b200/_lib/variants
for v in old_bin default old_bin default; do
  if [ $v = default ]; then unset MPTRAC_B200_LIBDIR; else export MPTRAC_B200_LIBDIR=$V/$v; fi
  MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --no-cpu --steps 12 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant':'$v', **{k: [round(v['ms_per_step'],4), round(v['ms_transport_only'],4)] for k, v in d['exchange'].items()}}))" | tee -a gpurun_out/sweep_bin3.jsonl
done
unset MPTRAC_B200_LIBDIR
MPB_BENCH_NO_SUSTAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c4g_r02r.csv \
  python bench.py --workload c4g --steps 6 --warmup 3 --no-cpu > gpurun_out/launches_c4g.log 2>&1; echo "launches c4g rc=$?"
MPB_BENCH_NO_SUSTAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5_r02r.csv \
  python bench.py --workload c5 --steps 6 --warmup 3 --no-cpu > gpurun_out/launches_c5.log 2>&1; echo "launches c5 rc=$?"
