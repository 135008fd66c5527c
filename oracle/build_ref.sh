#!/bin/bash
# TEST INFRASTRUCTURE -- builds the UNMODIFIED reference (slcs-jsc/mptrac) from the sources where
# they lie under /root/reference into oracle/_ref/ (git-ignored; travels to the GPU box with gpurun).
# Nothing under oracle/ is used by the product path; see DESIGN.md "Oracle".
#
# Recipe = SURVEY.md Appendix C: the reference's vendored dependency tarballs (libs/*.tar.bz2) are
# configured/built as static -fPIC archives, then src/mptrac.c and the handful of tools we use as
# fixtures are compiled directly with gcc using the reference's own gcc flags (src/Makefile:94-97).
# We do not run the reference's Makefile or libs/build.sh.
#
# usage: oracle/build_ref.sh [deps|ref|all]   (default all; idempotent, stamp files under _ref/)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${MPTRAC_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
B="$OUT/deps"
W="$OUT/build"
J="${JOBS:-$(nproc)}"
what="${1:-all}"

[ -d "$REF/src" ] || { echo "reference tree $REF not present: nothing to build"; exit 0; }
mkdir -p "$OUT/bin" "$OUT/lib" "$B" "$W/stub"
printf '#!/bin/sh\nexit 1\n' > "$W/stub/m4"; chmod +x "$W/stub/m4"   # netCDF configure only probes for m4

export CFLAGS="-fPIC -O2" CXXFLAGS="-fPIC -O2"

dep() {  # name, configure args...
  local name="$1"; shift
  [ -f "$B/.stamp_$name" ] && return 0
  echo "== building $name"
  rm -rf "$W/$name"; tar xjf "$REF/libs/$name.tar.bz2" -C "$W"
  local std="--disable-shared --enable-static"
  case "$name" in zlib-*) std="--static";; esac   # zlib ships its own (non-autoconf) configure
  ( cd "$W/$name" && PATH="$W/stub:$PATH" ./configure --prefix="$B" $std "$@" > configure.log 2>&1 \
      && make -j"$J" > make.log 2>&1 && make install > install.log 2>&1 ) \
      || { echo "FAILED: $name (see $W/$name/*.log)"; exit 1; }
  touch "$B/.stamp_$name"; rm -rf "$W/$name"
}

if [ "$what" = deps ] || [ "$what" = all ]; then
  dep gsl-2.7.1
  dep zlib-1.3.1
  dep szip-2.1.1
  dep hdf5-1.14.4-3 --with-zlib="$B" --with-szlib="$B" --enable-hl --disable-fortran --disable-cxx --disable-tests --disable-tools
  CPPFLAGS="-I$B/include" LDFLAGS="-L$B/lib" LIBS="-lhdf5_hl -lhdf5 -lsz -lz -ldl -lm" \
    dep netcdf-c-4.9.2 --disable-dap --disable-byterange --disable-nczarr --disable-libxml2 --disable-testsets --disable-utilities
fi

if [ "$what" = ref ] || [ "$what" = all ]; then
  SRC="$REF/src"
  DEFS="${MPTRAC_DEFINES:-}"
  # the reference's gcc flag set (src/Makefile:94-97), minus -Werror/-pedantic noise flags
  RFLAGS="-I$B/include $DEFS -DVERSION=\"oracle\" -O3 -g -DHAVE_INLINE -fno-common -fshort-enums -fopenmp"
  LIBS="-L$B/lib -lnetcdf -lhdf5_hl -lhdf5 -lsz -lz -lgsl -lgslcblas -ldl -lm"
  stamp="$OUT/.stamp_ref"
  # (the stamp records the -D dimensions: a different MPTRAC_DEFINES rebuilds everything that includes mptrac.h)
  if [ ! -f "$stamp" ] || [ "$SRC/mptrac.c" -nt "$stamp" ] || [ "$0" -nt "$stamp" ] || [ "$HERE/ref_harness.c" -nt "$stamp" ] \
     || [ "$HERE/ref_layout.c" -nt "$stamp" ] || [ "$(cat "$stamp" 2>/dev/null)" != "$DEFS" ]; then
    echo "== compiling reference library (static flavour, for timing + goldens)"
    gcc $RFLAGS -c "$SRC/mptrac.c" -o "$OUT/lib/mptrac.o"
    ar rcs "$OUT/lib/libmptrac.a" "$OUT/lib/mptrac.o"
    for t in trac atm_init atm_split atm_conv atm_dist atm_stat atm2grid wind sedi met_conv; do
      gcc $RFLAGS "$SRC/$t.c" "$OUT/lib/libmptrac.a" $LIBS -o "$OUT/bin/$t" &
    done
    wait
    echo "== compiling reference library (shared -fPIC flavour: the interposition boundary, SURVEY 8b)"
    gcc $RFLAGS -fPIC -shared -I"$SRC" "$SRC/mptrac.c" "$HERE/ref_layout.c" $LIBS -o "$OUT/lib/libmptrac.so"
    gcc $RFLAGS "$SRC/trac.c" -L"$OUT/lib" -lmptrac -Wl,-rpath,'$ORIGIN/../lib' $LIBS -o "$OUT/bin/trac_shared"
    echo "== compiling the in-memory harness around the reference translation unit"
    gcc $RFLAGS -DLOGLEV=0 -fPIC -fno-semantic-interposition -shared -I"$SRC" -I"$HERE" "$HERE/ref_harness.c" $LIBS -o "$OUT/lib/libref_harness.so"
    # headers needed to compile our shim + harness on a box that has no /root/reference are NOT copied:
    # everything that includes mptrac.h is compiled here, now.
    printf '%s' "$DEFS" > "$stamp"
  fi
fi
# test data of the reference's own tests that exercise this path (data files, not sources): they let the GPU box
# run the unmodified `trac` end to end through the shim.  oracle/_ref is git-ignored.
if [ ! -f "$OUT/data/.stamp6" ]; then
  mkdir -p "$OUT/data/dt_test.ref" "$OUT/data/coord_test.ref" "$OUT/data/trac_test.ref"
  cp "$REF"/tests/trac_test/data.ref/atm_init.tab "$REF"/tests/trac_test/data.ref/atm_pl_2011_*.tab "$REF"/tests/trac_test/data.ref/atm_ml_2011_*.tab "$OUT/data/trac_test.ref/"
  cp "$REF"/tests/data/era5ml_2011_06_0[5678]_00.nc "$OUT/data/"   # model-level met data of trac_test's atm_ml run
  # tests/interoper_test part 2: diabatic (zeta) transport on CLaMS-convention ERA-Interim data, ADVECT_VERT_COORD 1
  mkdir -p "$OUT/data/interoper_test.ref"
  cp "$REF"/tests/data/erai_vlr_1607010[06].nc "$OUT/data/"
  cp "$REF"/tests/interoper_test/data.ref/init/pos_glo_16070100.nc "$REF"/tests/interoper_test/data.ref/atm_2016_07_01_0[06]_00_00.tab "$OUT/data/interoper_test.ref/"
  cp "$REF"/tests/data/ei_2011_06_07_00.nc "$REF"/tests/data/ei_2011_06_08_00.nc "$OUT/data/"
  mkdir -p "$OUT/data/clim" && cp "$REF"/data/*.nc "$REF"/data/*.tab "$OUT/data/clim/"   # climatologies the chemistry modules read
  cp "$REF"/tests/data/ei_2011_06_05_00.nc "$REF"/tests/data/ei_2011_06_06_00.nc "$REF"/tests/data/era5_utm32_2025_05_01_0[012].nc "$OUT/data/"
  cp "$REF"/tests/dt_test/data.ref/atm_split.tab "$REF"/tests/dt_test/data.ref/atm_pl_*.tab "$OUT/data/dt_test.ref/"
  cp "$REF"/tests/coord_test/data.ref/atm_2025_05_01_*.tab "$OUT/data/coord_test.ref/"
  touch "$OUT/data/.stamp6"
fi
echo "oracle/_ref ready: $(ls "$OUT/bin" | tr '\n' ' ')"
