#!/bin/sh
# Build what can be built from the reference's own sources, in place.
# $1 = reference root.  Outputs only into oracle/_ref/.
set -e
REF="$1"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT"
# equipartition.{h,cpp} depend on libstdc++ only and compile unmodified.
g++ -O2 -march=native -fopenmp -fPIC -std=c++11 -I"$REF/src/ecckd" \
    -c "$REF/src/ecckd/equipartition.cpp" -o "$OUT/equipartition.o"
if [ -f "$HERE/adept_shim/adept_arrays.h" ] && [ -f "$HERE/ref_glue.cpp" ]; then
  sh "$HERE/build_ref_shim.sh" "$REF"
fi
if [ -f "$HERE/ref_mock/mock_files.cpp" ] && [ -f "$HERE/ref_tools_glue.cpp" ]; then
  sh "$HERE/build_ref_tools.sh" "$REF"
fi
