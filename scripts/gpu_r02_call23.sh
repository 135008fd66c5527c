#!/bin/bash
# Round 2, twenty-third GPU call (one GPU): model-level advection with timesteps / position checks folded into its launch
# (default) against the three-launch plan (MPTRAC_B200_NO_LEVEL_FOLD=1), same library, same box; then the whole GPU suite
# and the headline bench on this library.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/sweep_level_fold.jsonl
for v in nofold fold nofold fold; do
  if [ $v = fold ]; then unset MPTRAC_B200_NO_LEVEL_FOLD; else export MPTRAC_B200_NO_LEVEL_FOLD=1; fi
  MPB_BENCH_NO_SUSTAIN=1 timeout 400 python bench.py --workload c2ml --no-cpu --no-exchange --steps 48 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'workload':'c2ml','variant':'$v','ms_per_step':d['ms_per_step'],'frac':d['roofline']['frac'],'launches':d['gpu_launches']}))" | tee -a gpurun_out/sweep_level_fold.jsonl
done
unset MPTRAC_B200_NO_LEVEL_FOLD
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r02p.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_gpu_r02p.log
timeout 600 python bench.py > gpurun_out/bench_c2_r02p.json 2> gpurun_out/bench_c2_r02p.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_c2_r02p.json'))
print({k: d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'])
for k,v in d['exchange'].items(): print(k, v['ms_per_step'], v['ms_transport_only'])"
