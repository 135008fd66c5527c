#!/bin/bash
# Build launch-shape variants of the step kernel (here, no GPU needed) or time them (on the GPU box).
#   scripts/sweep_variants.sh build            -> mptrac_b200/_lib/variants/<name>/libmptrac_b200.so
#   scripts/sweep_variants.sh run [workload]   -> gpurun_out/sweep_<workload>.jsonl (one bench line per variant)
set -u
cd "$(dirname "$0")/.."
V=mptrac_b200/_lib/variants
# <flavour><block>m<min blocks per SM>; flavour f = fp32 cube kept across lookups, d = fp64 wind cube (MPB_CUBE_F64)
VARIANTS="${VARIANTS:-f128m3 f128m4 f128m5 f256m2 f64m8 d128m2 d128m3 d128m4 d64m6 d64m8}"
case "${1:-build}" in
  build)
    for v in $VARIANTS; do
      # <f|d><block>m<min blocks>  or  <f|d><block>r<register cap> (-maxrregcount, min blocks 1)
      fl=${v:0:1}; f64=0; [ "$fl" = d ] && f64=1; rest=${v:1}; cap=""
      if [[ "$rest" == *r* ]]; then b=${rest%r*}; m=1; cap="-maxrregcount=${rest#*r} -DMPB_NO_BOUNDS"; else b=${rest%m*}; m=${rest#*m}; fi
      mkdir -p $V/$v${TAG:-}
      nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fopenmp -shared \
        -DMPB_BLOCK=$b -DMPB_MINBLOCKS=$m -DMPB_CUBE_F64=$f64 $cap ${EXTRA:-} mptrac_b200/csrc/engine.cu -o $V/$v${TAG:-}/libmptrac_b200.so &
    done; wait; ls $V ;;
  run)
    WL=${2:-c2}; mkdir -p gpurun_out; : > gpurun_out/sweep_$WL.jsonl
    for v in $VARIANTS; do
      echo "== $v" >&2
      MPTRAC_B200_LIBDIR=$PWD/$V/$v MPB_BENCH_NO_SUSTAIN=1 timeout 300 python bench.py --workload $WL --no-cpu --steps 36 --warmup 3 2>/dev/null \
        | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(json.dumps({'variant':'$v','ms_per_step':d['ms_per_step'],'b2b_ms':d['back_to_back']['ms_per_step'],'value':d['value'],'frac':d['roofline']['frac']}))" | tee -a gpurun_out/sweep_$WL.jsonl
    done ;;
esac
