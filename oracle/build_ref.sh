#!/bin/bash
# Builds the UNMODIFIED reference (deal.II, read-only at /root/reference) into oracle/_ref/.
# TEST INFRASTRUCTURE ONLY: the result is the parity checker and the CPU baseline
# (bench.py --impl reference), never part of the product path.
#
# Recipe = SURVEY.md section 8(c) / BASELINE.md section 3 (serial, Release, bundled boost + Kokkos-Serial,
# no MPI / TBB / p4est: none of them is in this image).  LAPACK (needed by SolverCG's eigenvalue estimate, i.e.
# by PreconditionChebyshev and the multigrid of step-37) is the OpenBLAS inside this image's Python
# environment (opencv's libopenblasp with plain dstev_ ... symbols) when it is there, else OFF.  -march=x86-64-v4 instead of
# -march=native so the binaries run on the GPU box's host CPU too (AVX-512, VectorizedArray<double,8>).
#
# Outputs (all git-ignored):
#   oracle/_ref/build/     object files (also gpurun-ignored; can be deleted after the install)
#   oracle/_ref/install/   lib/libdeal_II.so + headers + cmake config
#   oracle/_ref/bin/       the drivers of oracle/ref_drivers/ (built by oracle/ref_drivers/build.sh)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF=${DEALII_SOURCE:-/root/reference}
OUT="$HERE/_ref"
JOBS=${JOBS:-8}
if [ ! -d "$REF" ]; then
  echo "build_ref.sh: $REF absent (GPU box?) - using the prebuilt oracle/_ref as is" >&2
  exit 0
fi
if [ -f "$OUT/install/lib/libdeal_II.so" ] && [ -z "${FORCE:-}" ]; then
  echo "build_ref.sh: $OUT/install/lib/libdeal_II.so exists (FORCE=1 to rebuild)"
  exit 0
fi
mkdir -p "$OUT/build" "$OUT/install"
cd "$OUT/build"
BLASDIR="${BLASDIR:-/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs}"
LAPACK_ARGS="-DDEAL_II_WITH_LAPACK=OFF"
if ls "$BLASDIR"/libopenblasp-*.so >/dev/null 2>&1; then
  OB="$(ls "$BLASDIR"/libopenblasp-*.so | head -1)"
  GF="$(ls "$BLASDIR"/libgfortran-*.so* | head -1)"
  QM="$(ls "$BLASDIR"/libquadmath-*.so* | head -1)"
  LAPACK_ARGS="-DDEAL_II_WITH_LAPACK=ON -DLAPACK_LIBRARIES=$OB;$GF;$QM -DBLAS_LIBRARIES=$OB"
fi
cmake -G Ninja "$REF" \
  -DCMAKE_BUILD_TYPE=Release \
  -DCMAKE_CXX_FLAGS="-march=x86-64-v4 -mprefer-vector-width=512" \
  -DDEAL_II_WITH_MPI=OFF "$LAPACK_ARGS" -DDEAL_II_WITH_TBB=OFF -DDEAL_II_WITH_P4EST=OFF \
  -DDEAL_II_COMPONENT_EXAMPLES=OFF -DDEAL_II_ALLOW_AUTODETECTION=OFF \
  -DDEAL_II_FORCE_BUNDLED_BOOST=ON -DDEAL_II_FORCE_BUNDLED_KOKKOS=ON \
  -DCMAKE_INSTALL_PREFIX="$OUT/install" > "$OUT/configure.log" 2>&1
ninja -j"$JOBS" install > "$OUT/build.log" 2>&1
echo "build_ref.sh: installed into $OUT/install"
