#!/bin/bash
# Round 2, twenty-first GPU call (one GPU): model-level advection -- neighbouring intervals tried before the bisection
# (default build) against the build without (variants/nonear), and block sizes 64 / 32 (variants/b64, b32), same box.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_shim_trac.py -m gpu -q -x > gpurun_out/pytest_near.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/pytest_near.log
: > gpurun_out/sweep_level_near.jsonl
V=$PWD/mptrac_b200/_lib/variants
for v in nonear default b64 b32 nonear default b64 b32; do
  if [ $v = default ]; then unset MPTRAC_B200_LIBDIR; else export MPTRAC_B200_LIBDIR=$V/$v; fi
  MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --workload c2ml --no-cpu --no-exchange --steps 48 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'workload':'c2ml','variant':'$v','ms_per_step':d['ms_per_step'],'frac':d['roofline']['frac']}))" | tee -a gpurun_out/sweep_level_near.jsonl
done
