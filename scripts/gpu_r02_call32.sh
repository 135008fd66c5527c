#!/bin/bash
# Round 2, thirty-third GPU call (one GPU): the binning kernel with its blocks scattered over the parcel array (lane-private partial sums, no
# scan in crowded rounds) against the previous library, launch lists including cell sorts; parity of the grid / mixing tests.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_dist.py -m gpu -q -x -k "grid or mix or dist or team or peers or process or allreduce" > gpurun_out/pytest_bin.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/pytest_bin.log
V=$PWD/mptrac_b200/_lib/variants
for v in old_bin default; do
  if [ $v = default ]; then unset MPTRAC_B200_LIBDIR; else export MPTRAC_B200_LIBDIR=$V/$v; fi
  MPB_BENCH_NO_SUSTAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c4g_sorted_$v.csv \
    python bench.py --workload c4g --steps 24 --warmup 3 --no-cpu --no-exchange > gpurun_out/launches_c4g_$v.log 2>&1; echo "launches $v rc=$?"
done
unset MPTRAC_B200_LIBDIR
for v in old_bin default old_bin default; do
  if [ $v = default ]; then unset MPTRAC_B200_LIBDIR; else export MPTRAC_B200_LIBDIR=$V/$v; fi
  MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --no-cpu --steps 12 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant':'$v', **{k: [round(v['ms_per_step'],4), round(v['ms_transport_only'],4)] for k, v in d['exchange'].items()}}))" | tee -a gpurun_out/sweep_bin5.jsonl
done
