#!/bin/bash
# The reference's UNMODIFIED `trac` driver at BASELINE C2 scale (~1 M parcels, 360x181x61 synthetic met from the
# reference's own `wind` tool, RK4, 72 steps of 300 s), once as the reference runs it (OpenMP on the host cores) and once
# with libmptrac_b200_shim.so pre-loaded; prints both timer summaries and the largest difference of the final parcel
# positions.  Runs on the GPU box (needs oracle/_ref).   usage: scripts/trac_dropin_bench.sh [meteo]
set -eu
cd "$(dirname "$0")/.."
R=$PWD/oracle/_ref/bin; SHIM=$PWD/mptrac_b200/_lib/libmptrac_b200_shim.so
W=$(mktemp -d); mkdir -p $W/data gpurun_out
# (quantity idx = parcel index: SORT_DT reorders the parcels, the comparison matches them by index)
QNT=$'NQ = 1\nQNT_NAME[0] = idx'; MDO="MET_DT_OUT = 0"; NQ=1
if [ "${1:-}" = meteo ]; then QNT=$'NQ = 5\nQNT_NAME[0] = idx\nQNT_NAME[1] = t\nQNT_NAME[2] = u\nQNT_NAME[3] = v\nQNT_NAME[4] = w'; MDO=""; NQ=5; fi
cat > $W/data/trac.ctl <<CTL
$QNT
METBASE = $W/data/wind
DT_MET = 21600
DT_MOD = 300
T_STOP = 21600
ADVECT = 4
DIFFUSION = 0
$MDO
SORT_DT = 3600
ATM_DT_OUT = 21600
ATM_TYPE = 1
ATM_TYPE_OUT = 1
MET_CAPE = 0
MET_PBL = 0
CTL
echo $W/data > $W/dirlist
export LANG=C LC_ALL=C OMP_NUM_THREADS=$(nproc)
for t in 0 21600; do
  $R/wind $W/data/trac.ctl $W/data/wind WIND_T0 $t WIND_NX 360 WIND_NY 181 WIND_NZ 61 WIND_Z0 0 WIND_Z1 60 WIND_ALPHA 45 \
     WIND_U0 38.59 WIND_U1 60 WIND_W0 0.01 WIND_TEMP0 280 WIND_TEMP1 220 > $W/wind_$t.log 2>&1
done
$R/atm_init $W/data/trac.ctl $W/data/atm_init.tab INIT_T0 0 INIT_T1 0 INIT_LON0 -180 INIT_LON1 179 INIT_DLON 1 \
   INIT_LAT0 -89.5 INIT_LAT1 89.5 INIT_DLAT 1 INIT_Z0 2 INIT_Z1 30 INIT_DZ 2 > $W/init.log 2>&1
grep -i "number of\|np =" $W/init.log | tail -2 || true
run() {  # name, binary, preload, [VAR=value ...]
  local name=$1 bin=$2 pre=$3; shift 3
  mkdir -p $W/$name; cp $W/data/trac.ctl $W/data/atm_init.tab $W/$name/; ln -sf $W/data/wind_*.nc $W/$name/ 2>/dev/null || true
  sed -i "s|METBASE = .*|METBASE = $W/data/wind|" $W/$name/trac.ctl
  echo $W/$name > $W/dirlist_$name
  local t0=$(date +%s.%N)
  ( for kv in "$@"; do export "$kv"; done; if [ -n "$pre" ]; then export LD_PRELOAD=$pre MPTRAC_B200_VERBOSE=1; fi
    $bin $W/dirlist_$name trac.ctl atm_init.tab ATM_BASENAME atm > $W/$name.log 2>&1 ) || { echo "$name FAILED"; tail -5 $W/$name.log; }
  set -- $name
  local t1=$(date +%s.%N)
  echo "== $1: wall $(python -c "print(f'{$t1 - $t0:.2f}')") s"
  grep -E "SIZE_NP|TIMER_GROUP_PHYSICS|TIMER_GROUP_INPUT|TIMER_GROUP_MEMORY|TIMER_MODULE_ADVECT|TIMER_MODULE_B200_STEP|TIMER_MODULE_METEO|TIMER_MODULE_SORT|TIMER_TOTAL|kernel launches" $W/$1.log || tail -5 $W/$1.log
}
run cpu $R/trac ""
run gpu_cold $R/trac_shared $SHIM          # first CUDA process on a fresh box: the driver itself is still being paged in
run gpu $R/trac_shared $SHIM
run gpu_no_warmup $R/trac_shared $SHIM MPTRAC_B200_NO_WARMUP=1   # context created when the parcels arrive, not while trac reads
# the strict arithmetic flavour (-fmad=false, correctly rounded quotients) behind the same shim
mkdir -p $W/strictlib; cp $SHIM $W/strictlib/; cp $PWD/mptrac_b200/_lib/libmptrac_b200_strict.so $W/strictlib/libmptrac_b200.so
run gpu_strict $R/trac_shared $W/strictlib/libmptrac_b200_shim.so
compare() {
python - <<PY
import numpy as np, glob
def rd(f):
    a = np.fromfile(f, dtype=np.uint8)
    # ATM_TYPE_OUT 1: int version, int np, then time, p, lon, lat, q[nq] as doubles (src/mptrac.c:12872-12918)
    hdr = np.frombuffer(a[:8].tobytes(), dtype=np.int32); n = int(hdr[1])
    d = np.frombuffer(a[8:8 + 8 * (4 + $NQ) * n].tobytes(), dtype=np.float64).reshape(4 + $NQ, n)
    return n, d[:, np.argsort(d[4])]          # ordered by parcel index
fc = sorted(glob.glob("$W/cpu/atm_2*.bin"))[-1]; fg = sorted(glob.glob("$W/$1/atm_2*.bin"))[-1]
print("-- $1 against the reference's CPU run")
n, c = rd(fc); m, g = rd(fg)
assert n == m and np.array_equal(c[4], g[4])
dlon = ((c[2] - g[2] + 180.0) % 360.0 - 180.0) * np.cos(np.deg2rad(c[3]))
d = np.hypot(dlon, c[3] - g[3])
q = np.quantile(d, [0.5, 0.99, 0.999, 1.0])
bad = d > 1e-8
print(f"parcels {n}: position difference [deg] median {q[0]:.2e}, 99% {q[1]:.2e}, 99.9% {q[2]:.2e}, max {q[3]:.2e}; "
      f"max rel dp {np.max(np.abs(c[1]-g[1])/c[1]):.2e}")
print(f"  {bad.sum()} parcels differ by more than 1e-8 deg; they end at |lat| >= {np.abs(c[3][bad]).min() if bad.any() else 0:.2f} deg "
      f"(the trajectories that pass within 0.001 deg of a pole, where DX2DEG drops to zero: src/mptrac.h:904)")
if $NQ > 1:
    ok = ~bad
    print(f"  quantities (parcels that agree in position): max rel-to-scale {np.max(np.abs(c[5:, ok]-g[5:, ok]) / np.max(np.abs(c[5:]), axis=1, keepdims=True)):.2e}")
PY
}
compare gpu
compare gpu_strict
NG=$(nvidia-smi -L 2>/dev/null | wc -l)
if [ "$NG" -gt 1 ]; then    # the same simulation cut over all GPUs of the box behind trac's one host thread
  run gpu_team$NG $R/trac_shared $SHIM MPTRAC_B200_DEVICES=0-$((NG-1))
  compare gpu_team$NG
fi
cp $W/cpu.log gpurun_out/trac_dropin_cpu.log; cp $W/gpu.log gpurun_out/trac_dropin_gpu.log; cp $W/gpu_strict.log gpurun_out/trac_dropin_gpu_strict.log
rm -rf $W
