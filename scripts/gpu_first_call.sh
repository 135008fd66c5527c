#!/bin/bash
# First GPU call of a round (under gpurun, one GPU, ~4 min of box time): the gpu suite, the headline bench line and its launch
# list.  Everything lands in gpurun_out/.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- scripts/gpu_first_call.sh
# Optional, built beforehand where nvcc is (no GPU needed): the record-cache flavour of the model-level advection,
#   VARIANTS=f128m4 TAG=_lc EXTRA=-DMPB_LEVEL_CACHE=1 scripts/sweep_variants.sh build
# -- if mptrac_b200/_lib/variants/f128m4_lc exists, its parity tests and its c2ml bench line run next to the shipped one.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -3 gpurun_out/pytest_gpu.log
scripts/gpu_check.sh bench launches
WL=c2ml scripts/gpu_check.sh bench
LC=$PWD/mptrac_b200/_lib/variants/f128m4_lc
if [ -f "$LC/libmptrac_b200.so" ]; then
  MPTRAC_B200_LIBDIR=$LC timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k model_level_advection_vs_oracle 2>&1 | tail -2
  MPTRAC_B200_LIBDIR=$LC MPB_BENCH_NO_SUSTAIN=1 timeout 300 python bench.py --workload c2ml --no-cpu --steps 36 --warmup 3 > gpurun_out/bench_c2ml_level_cache.json 2> gpurun_out/bench_c2ml_level_cache.err
  python -c "import json; d=json.load(open('gpurun_out/bench_c2ml_level_cache.json')); print('c2ml with record cache: ms/step', d['ms_per_step'], 'value', d['value'])"
fi
