#!/bin/sh
# Compile the reference's own hot-path translation units, unmodified and where they lie
# under $1/src/ecckd, against the Adept API shim (oracle/adept_shim), and link them with
# the C-ABI glue into oracle/_ref/libecckd_ref.so.  Nothing is copied out of the reference.
# Flags follow configure_ecmwf.sh:8 (-O2 -march=native); OpenMP is left off because the
# reference throws from inside `omp parallel for` on its error paths (std::terminate), and
# the glue wants those paths to return an exit code instead; this changes no arithmetic.
set -e
REF="$1"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT"
UNITS="average_optical_depth planck_function radiative_transfer_lw radiative_transfer_sw single_gas_data calc_cost_function_lw calc_cost_function_sw"
FLAGS="-O2 -march=native -fPIC -std=c++11 -Wno-unknown-pragmas -fvisibility=hidden -I$HERE/adept_shim -I$REF/src/ecckd -I$REF/src/include"
OBJS=""
for u in $UNITS; do
  # average_optical_depth.cpp relies on its caller having included <adept_arrays.h>-dependent
  # declarations through constants.h; every unit compiles as is.
  g++ $FLAGS -fvisibility=default -c "$REF/src/ecckd/$u.cpp" -o "$OUT/$u.o"
  OBJS="$OBJS $OUT/$u.o"
done
g++ $FLAGS -c "$HERE/ref_glue.cpp" -o "$OUT/ref_glue.o"
g++ -shared -o "$OUT/libecckd_ref.so" "$OUT/ref_glue.o" $OBJS
rm -f $OBJS "$OUT/ref_glue.o"
echo "built $OUT/libecckd_ref.so"
