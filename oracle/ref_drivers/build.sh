#!/bin/bash
# Builds the drivers that link the reference library of oracle/build_ref.sh (TEST INFRASTRUCTURE:
# parity checker and CPU baseline only) into oracle/_ref/bin/.  Plain g++ on the files of this
# directory; the reference's own headers/library are used where oracle/build_ref.sh installed them.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="$HERE/../_ref/install"
BIN="$HERE/../_ref/bin"
if [ ! -f "$REF/lib/libdeal_II.so" ] || [ ! -d "$REF/include/deal.II" ]; then
  echo "ref_drivers/build.sh: no reference install with headers at $REF (run oracle/build_ref.sh where /root/reference exists)" >&2
  exit 0
fi
mkdir -p "$BIN"
CXX=${CXX:-g++}
FLAGS="-std=c++17 -O2 -fopenmp-simd -march=x86-64-v4 -mprefer-vector-width=512 -Wno-unused-parameter -Wno-deprecated-declarations"
INC="-I$REF/include -I$REF/include/deal.II/bundled"
LINK="-L$REF/lib -ldeal_II -Wl,-rpath,\$ORIGIN/../install/lib -rdynamic -ldl -lpthread"
# the reference was configured with LAPACK = the OpenBLAS that ships inside this image's Python environment
# (oracle/build_ref.sh); that library needs its libgfortran, which has no rpath of its own: link it into the
# drivers so that it is loaded before OpenBLAS asks for it
BLASDIR="${BLASDIR:-/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs}"
if readelf -d "$REF/lib/libdeal_II.so" | grep -q openblas; then
  LINK="$LINK -Wl,--no-as-needed $(ls $BLASDIR/libgfortran-*.so* | head -1) $(ls $BLASDIR/libquadmath-*.so* | head -1) -Wl,--as-needed -Wl,-rpath,$BLASDIR"
fi
pids=()
for deg in ${DEGREES:-1 2 3 4 5 6 7 8}; do
  if [ ! -x "$BIN/ref_dump_q$deg" ] || [ "$HERE/ref_dump.cc" -nt "$BIN/ref_dump_q$deg" ]; then
    ( $CXX $FLAGS -DREF_DEGREE=$deg $INC "$HERE/ref_dump.cc" -o "$BIN/ref_dump_q$deg" $LINK ) &
    pids+=($!)
  fi
done
# over-integrated variants of ref_dump (n_q_points_1d = degree + 1 + extra) for tests/golden/ref_nq
for spec in ${NQ_SPECS:-2:1 3:1}; do
  deg=${spec%%:*}; extra=${spec##*:}
  if [ ! -x "$BIN/ref_dump_q${deg}_nq${extra}" ] || [ "$HERE/ref_dump.cc" -nt "$BIN/ref_dump_q${deg}_nq${extra}" ]; then
    ( $CXX $FLAGS -DREF_DEGREE=$deg -DREF_NQ_EXTRA=$extra $INC "$HERE/ref_dump.cc" -o "$BIN/ref_dump_q${deg}_nq${extra}" $LINK ) &
    pids+=($!)
  fi
done
for deg in ${GMG_DEGREES:-1 2 3 4}; do
  if [ ! -x "$BIN/ref_gmg_q$deg" ] || [ "$HERE/ref_gmg.cc" -nt "$BIN/ref_gmg_q$deg" ]; then
    ( $CXX $FLAGS -DREF_DEGREE=$deg $INC "$HERE/ref_gmg.cc" -o "$BIN/ref_gmg_q$deg" $LINK ) &
    pids+=($!)
  fi
done
if [ ! -x "$BIN/ref_bench" ] || [ "$HERE/ref_bench.cc" -nt "$BIN/ref_bench" ]; then
  ( $CXX $FLAGS $INC "$HERE/ref_bench.cc" -o "$BIN/ref_bench" $LINK ) &
  pids+=($!)
fi
# the product-side deal.II adapter example (include/b200mf_dealii.hpp + examples/step64_dealii.cc):
# deal.II host code driving libb200mf.so; it can only be compiled where deal.II's headers are
ROOT="$HERE/../.."
if [ -f "$ROOT/dealii_b200/libb200mf.so" ]; then
  for ex in step64 step37; do
    if [ ! -x "$BIN/${ex}_b200" ] || [ "$ROOT/examples/${ex}_dealii.cc" -nt "$BIN/${ex}_b200" ] || [ "$ROOT/include/b200mf_dealii.hpp" -nt "$BIN/${ex}_b200" ] \
       || [ "$ROOT/include/b200mf_portable.hpp" -nt "$BIN/${ex}_b200" ] || [ "$ROOT/include/b200mf.h" -nt "$BIN/${ex}_b200" ]; then
      ( $CXX $FLAGS $INC -I"$ROOT/include" -I/usr/local/cuda/include "$ROOT/examples/${ex}_dealii.cc" -o "$BIN/${ex}_b200" \
          $LINK -L"$ROOT/dealii_b200" -lb200mf -Wl,-rpath,\$ORIGIN/../../../dealii_b200 -L/usr/local/cuda/lib64 -lcudart ) &
      pids+=($!)
    fi
  done
fi
rc=0
for p in "${pids[@]:-}"; do [ -z "$p" ] || wait "$p" || rc=1; done
exit $rc
