#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests, the bench line, the ncu launch list and one full capture of
# the step kernel.  Everything lands in gpurun_out/.   usage: scripts/gpu_check.sh [tests] [bench] [launches] [ncu]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what="${*:-tests bench launches ncu}"
WL="${WL:-c2}"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
for w in $what; do
  case $w in
    tests)    timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log ;;
    bench)    timeout 420 python bench.py --workload $WL > gpurun_out/bench_$WL.json 2> gpurun_out/bench_$WL.err; echo "bench rc=$?"; cat gpurun_out/bench_$WL.json ;;
    benchref) timeout 900 python bench.py --impl reference --workload $WL --steps 5 --warmup 1 > gpurun_out/bench_ref_$WL.json 2> gpurun_out/bench_ref_$WL.err; cat gpurun_out/bench_ref_$WL.json ;;
    launches) MPB_BENCH_NO_SUSTAIN=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$WL.csv \
                python bench.py --workload $WL --steps 24 --warmup 3 --no-cpu > gpurun_out/launches_$WL.log 2>&1; echo "launches rc=$?" ;;
    ncu)      MPB_BENCH_NO_SUSTAIN=1 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s ${NCU_SKIP:-14} -c 2 -f -o gpurun_out/prof_$WL \
                python bench.py --workload $WL --steps 24 --warmup 3 --no-cpu > gpurun_out/ncu_$WL.log 2>&1; echo "ncu rc=$?" ;;
  esac
done
