#!/usr/bin/env bash
# Build the REAL reference (from $REF_DIR, default /root/reference) into oracle/_ref/.
#
# TEST / BASELINE INFRASTRUCTURE ONLY. Outputs (git-ignored, but they travel to the GPU box):
#   oracle/_ref/ref_driver        op-script runner over the reference API (x86-64-v3 build)
#   oracle/_ref/ref_driver_v4     same, x86-64-v4 (AVX-512) build; picked at run time if the host has avx512f
#   oracle/_ref/ref_tests         the reference's own Catch2 suite (only with --with-tests, ~1 min to compile)
#
# The reference needs MPI, which does not exist in this image: oracle/mpi_shim/mpi.h stands in
# (fork + shared-memory rings, SHIM_NP=P replaces `mpirun -np P`).
# The reference as shipped has a one-token bug in setBit (src/bit_maths.hpp:66) that makes
# manyTargGate/pauli*/krausMap/partialTrace garbage; it is patched in a throw-away copy under $TMPDIR --
# reference sources are never copied into this repo.  Flags follow compile.sh:2 except -march (native
# is not portable from this container to the GPU box's host CPU).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF_DIR="${REF_DIR:-/root/reference}"
out="$here/_ref"
mkdir -p "$out"

if [ ! -d "$REF_DIR/src" ]; then
    echo "build_ref: $REF_DIR not present; keeping prebuilt files in $out (if any)"
    exit 0
fi

tmp="$(mktemp -d)"
trap 'rm -rf "$tmp"' EXIT
cp -r "$REF_DIR/src" "$tmp/src"
# the F1 patch: keep every bit EXCEPT the addressed one, then OR in the new value
sed -i 's/return (number & (~comp)) | mask;/return (number \& comp) | mask;/' "$tmp/src/bit_maths.hpp"
grep -q 'return (number & comp) | mask;' "$tmp/src/bit_maths.hpp" || { echo "build_ref: setBit patch did not apply"; exit 1; }

conf="-std=c++17 -O3 -fopenmp"
inc="-I$here/mpi_shim -I$tmp/src"
echo "build_ref: ref_driver (v3)"
g++ $conf -march=x86-64-v3 $inc "$here/ref_driver.cpp" -o "$out/ref_driver"
echo "build_ref: ref_driver (v4)"
g++ $conf -march=x86-64-v4 $inc "$here/ref_driver.cpp" -o "$out/ref_driver_v4"

if [ "${1:-}" = "--with-tests" ]; then
    echo "build_ref: ref_tests (reference's own Catch2 suite)"
    g++ $conf -march=x86-64-v3 $inc -I"$REF_DIR/tests" -I"$REF_DIR/catch" \
        "$REF_DIR/tests/tests.cpp" "$REF_DIR/catch/catch_amalgamated.cpp" -o "$out/ref_tests"
fi
echo "build_ref: done -> $out"
