#!/bin/bash
# Builds the UNMODIFIED reference (deal.II, read-only at /root/reference) into oracle/_ref/.
# TEST INFRASTRUCTURE ONLY: the result is the parity checker and the CPU baseline
# (bench.py --impl reference), never part of the product path.
#
# Recipe = SURVEY.md section 8(c) / BASELINE.md section 3 (serial, Release, bundled boost + Kokkos-Serial,
# no MPI / LAPACK / TBB / p4est: none of them is in this image).  -march=x86-64-v4 instead of
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
cmake -G Ninja "$REF" \
  -DCMAKE_BUILD_TYPE=Release \
  -DCMAKE_CXX_FLAGS="-march=x86-64-v4 -mprefer-vector-width=512" \
  -DDEAL_II_WITH_MPI=OFF -DDEAL_II_WITH_LAPACK=OFF -DDEAL_II_WITH_TBB=OFF -DDEAL_II_WITH_P4EST=OFF \
  -DDEAL_II_COMPONENT_EXAMPLES=OFF -DDEAL_II_ALLOW_AUTODETECTION=OFF \
  -DDEAL_II_FORCE_BUNDLED_BOOST=ON -DDEAL_II_FORCE_BUNDLED_KOKKOS=ON \
  -DCMAKE_INSTALL_PREFIX="$OUT/install" > "$OUT/configure.log" 2>&1
ninja -j"$JOBS" install > "$OUT/build.log" 2>&1
echo "build_ref.sh: installed into $OUT/install"
