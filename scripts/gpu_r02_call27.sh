#!/bin/bash
# Round 2, twenty-seventh GPU call (one GPU): binning with the grid's cell sizes precomputed and counts taken from run lengths
# (default) against the previous library (variants/old_bin), same box: exchange records of c4g and c5; parity first.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_modules.py tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/pytest_bin.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/pytest_bin.log
: > gpurun_out/sweep_bin.jsonl
V=$PWD/mptrac_b200/_lib/variants
for v in old_bin default old_bin default; do
  if [ $v = default ]; then unset MPTRAC_B200_LIBDIR; else export MPTRAC_B200_LIBDIR=$V/$v; fi
  MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --no-cpu --steps 12 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant':'$v', **{k: [round(v['ms_per_step'],4), round(v['ms_transport_only'],4)] for k, v in d['exchange'].items()}}))" | tee -a gpurun_out/sweep_bin.jsonl
done
unset MPTRAC_B200_LIBDIR
MPB_BENCH_NO_SUSTAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c4g_r02r.csv \
  python bench.py --workload c4g --steps 6 --warmup 3 --no-cpu > gpurun_out/launches_c4g.log 2>&1; echo "launches c4g rc=$?"
grep -c grid_bin_kernel gpurun_out/launches_c4g_r02r.csv
