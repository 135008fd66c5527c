#!/bin/bash
# Round 2, third GPU call: the lane-per-coordinate step kernel -- bit-identity with the classic kernel, the parity suites with
# it as the default, timing of its occupancy variants against the classic kernel, one full ncu capture.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_quad.py tests/test_gpu_parity.py tests/test_gpu_scale.py tests/test_gpu_modules.py -m gpu -q -x > gpurun_out/pytest_quad.log 2>&1; echo "quad + parity rc=$?"; tail -15 gpurun_out/pytest_quad.log
: > gpurun_out/sweep_quad.jsonl
run() {  # name, env...
  local name=$1; shift
  env "$@" MPB_BENCH_NO_SUSTAIN=1 timeout 300 python bench.py --workload c2 --no-cpu --no-exchange --steps 48 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant':'$name','ms_per_step':d['ms_per_step'],'b2b_ms':d['back_to_back']['ms_per_step'],'e2e_ms':d['e2e']['ms_per_step'],'frac':d['roofline']['frac']}))" | tee -a gpurun_out/sweep_quad.jsonl
}
run classic MPTRAC_B200_STEP=classic
run quad_q6 MPTRAC_B200_STEP=quad
for v in q4 q5 q8; do run quad_$v MPTRAC_B200_LIBDIR=$PWD/mptrac_b200/_lib/variants/$v; done
MPB_BENCH_NO_SUSTAIN=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:quad_step_kernel -s 10 -c 1 -f -o gpurun_out/prof_quad_c2 \
  python bench.py --workload c2 --steps 24 --warmup 3 --no-cpu --no-exchange > gpurun_out/ncu_quad.log 2>&1; echo "ncu rc=$?"
