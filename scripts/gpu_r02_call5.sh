#!/bin/bash
# Round 2, fifth GPU call (one GPU): TMA-tile variant -- parity and timing on the sparse (c2) and the dense (c4) workload --,
# quad variant's tests, launch list of the configs[3] step with its gridded output, the unmodified trac drop-in at C2 scale.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_quad.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_tile.log 2>&1; echo "tile/quad/parity tests rc=$?"; tail -4 gpurun_out/pytest_tile.log
: > gpurun_out/sweep_tile.jsonl
run() {  # name workload env...
  local name=$1 wl=$2; shift 2
  env "$@" MPB_BENCH_NO_SUSTAIN=1 timeout 300 python bench.py --workload $wl --no-cpu --no-exchange --steps 36 --warmup 3 2>/dev/null \
    | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant':'$name','workload':'$wl','ms_per_step':d['ms_per_step'],'b2b_ms':d['back_to_back']['ms_per_step'],'frac':d['roofline']['frac']}))" | tee -a gpurun_out/sweep_tile.jsonl
}
run classic c2 MPTRAC_B200_STEP=classic
run tile_3x4x32 c2 MPTRAC_B200_STEP=tile
run tile_2x6x60 c2 MPTRAC_B200_STEP=tile MPTRAC_B200_TILE=2,6,60
run tile_3x8x60 c2 MPTRAC_B200_STEP=tile MPTRAC_B200_TILE=3,8,60
run tile_3x4x32_sort1200 c2 MPTRAC_B200_STEP=tile MPB_BENCH_SORT_DT=1200
run classic_sort1200 c2 MPTRAC_B200_STEP=classic MPB_BENCH_SORT_DT=1200
run classic c4 MPTRAC_B200_STEP=classic
run tile_3x4x32 c4 MPTRAC_B200_STEP=tile
run tile_3x3x16 c4 MPTRAC_B200_STEP=tile MPTRAC_B200_TILE=3,3,16
run tile_4x6x32 c4 MPTRAC_B200_STEP=tile MPTRAC_B200_TILE=4,6,32
MPTRAC_B200_STEP=tile MPB_BENCH_NO_SUSTAIN=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_step_kernel -s 10 -c 1 -f -o gpurun_out/prof_tile_c2 \
  python bench.py --workload c2 --steps 24 --warmup 3 --no-cpu --no-exchange > gpurun_out/ncu_tile.log 2>&1; echo "ncu tile rc=$?"
MPB_BENCH_NO_SUSTAIN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c4g.csv \
  python bench.py --workload c4g --steps 6 --warmup 3 --no-cpu > gpurun_out/launches_c4g.log 2>&1; echo "launches c4g rc=$?"
timeout 900 scripts/trac_dropin_bench.sh > gpurun_out/trac_dropin.txt 2>&1; echo "dropin rc=$?"; tail -25 gpurun_out/trac_dropin.txt; grep "mptrac_b200:" gpurun_out/trac_dropin_gpu.log | head -12
