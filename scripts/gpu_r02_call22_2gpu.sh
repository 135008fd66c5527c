#!/bin/bash
# Round 2, twenty-second GPU call (2 GPUs): where the mixing exchange spends its time -- phase times of module_mixing
# (MPTRAC_B200_TRACE_MIXING: CUDA events around prepare / route / barrier / serve / barrier / apply) on 1 and 2 ranks.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for n in 1 2; do
  if [ $n = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517"; fi
  MPTRAC_B200_TRACE_MIXING=1 MPB_BENCH_NO_SUSTAIN=1 timeout 600 $L bench.py --gpus $n --steps 6 --warmup 3 --no-cpu > gpurun_out/trace_mix_${n}gpu.json 2> gpurun_out/trace_mix_${n}gpu.err
  echo "n=$n rc=$?"; grep "mixing trace" gpurun_out/trace_mix_${n}gpu.err
  python -c "import json; d=json.load(open('gpurun_out/trace_mix_${n}gpu.json')); print(json.dumps(d['exchange']['c5'], indent=0)[:600])"
done
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_scale.py tests/test_gpu_parity.py -m gpu -q -x -k "mix or dist or team or peers or allreduce or process" 2>&1 | tail -3
