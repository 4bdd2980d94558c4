#!/bin/sh
# Compile the reference's own TOOLS (run_ckd, create_look_up_table, reorder_spectrum, find_g_points),
# unmodified and where they lie under $1/src, into oracle/_ref/libecckd_ref_tools.so:
#   * array maths against the Adept API shim (oracle/adept_shim);
#   * configuration files through the reference's own cfg engine (DataFileEngineCfg.cpp, readconfig.c);
#   * NetCDF files through the in-memory stand-in oracle/ref_mock/mock_files.cpp (NetCDF is not
#     installed here);  netcdf.h / config.h stubs in oracle/ref_mock.
# Every tool is compiled with its `main` intact (a function called main returns 0 when control
# reaches its end; renamed in the source it would not) and the symbol is renamed afterwards
# (objcopy --redefine-sym main=ref_main_<tool>); the mains are called through oracle/ref_tools_glue.cpp.  Nothing is copied out of the reference.  Outputs only into oracle/_ref/.
set -e
REF="$1"
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
mkdir -p "$OUT/tools_obj"
FLAGS="-O2 -march=native -fPIC -std=c++11 -Wno-unknown-pragmas -DHAVE_CONFIG_H -I$HERE/ref_mock -I$HERE/adept_shim -I$REF/src/ecckd -I$REF/src/include"
OBJS=""
for u in average_optical_depth planck_function radiative_transfer_lw radiative_transfer_sw single_gas_data \
         calc_cost_function_lw calc_cost_function_sw ckd_model read_spectrum read_merged_spectrum write_order; do
  g++ $FLAGS -fopenmp -c "$REF/src/ecckd/$u.cpp" -o "$OUT/tools_obj/$u.o" &
  OBJS="$OBJS $OUT/tools_obj/$u.o"
done
# equipartition.cpp: the partition search is a discrete iteration whose path depends on the last bit of every
# bound, so the oracle build evaluates its formulas in strict IEEE double arithmetic (no a*b+c contraction into
# FMAs, which -march=native would otherwise allow) - the arithmetic ecckd_partition_search_* is held to, bit for bit.
g++ $FLAGS -fopenmp -ffp-contract=off -c "$REF/src/ecckd/equipartition.cpp" -o "$OUT/tools_obj/equipartition.o" &
OBJS="$OBJS $OUT/tools_obj/equipartition.o"
for t in run_ckd create_look_up_table reorder_spectrum; do
  (g++ $FLAGS -c "$REF/src/ecckd/$t.cpp" -o "$OUT/tools_obj/main_$t.o" &&
   objcopy --redefine-sym main=ref_main_$t "$OUT/tools_obj/main_$t.o") &
  OBJS="$OBJS $OUT/tools_obj/main_$t.o"
done
for u in Logging Error file_manager DataFileEngineCfg; do
  g++ $FLAGS -c "$REF/src/tools/$u.cpp" -o "$OUT/tools_obj/$u.o" &
  OBJS="$OBJS $OUT/tools_obj/$u.o"
done
gcc -O2 -fPIC -I"$HERE/ref_mock" -I"$REF/src/include" -c "$REF/src/tools/readconfig.c" -o "$OUT/tools_obj/readconfig.o" &
OBJS="$OBJS $OUT/tools_obj/readconfig.o"
g++ $FLAGS -c "$HERE/ref_mock/mock_files.cpp" -o "$OUT/tools_obj/mock_files.o" &
g++ $FLAGS -c "$HERE/ref_tools_glue.cpp" -o "$OUT/tools_obj/ref_tools_glue.o" &
# find_g_points.cpp is compiled through the glue that includes it (its fit functions and the
# CkdEquipartition class have no header)
(g++ $FLAGS -fopenmp -c "$HERE/ref_tools_fgp_glue.cpp" -o "$OUT/tools_obj/ref_tools_fgp_glue.o" &&
 objcopy --redefine-sym main=ref_main_find_g_points "$OUT/tools_obj/ref_tools_fgp_glue.o") &
wait
g++ -shared -fopenmp -o "$OUT/libecckd_ref_tools.so" "$OUT/tools_obj/ref_tools_glue.o" "$OUT/tools_obj/ref_tools_fgp_glue.o" "$OUT/tools_obj/mock_files.o" $OBJS
echo "built $OUT/libecckd_ref_tools.so"

# The same tools as a CONSUMER of the C ABI: integration/ecckd_gpu_bindings.cpp takes the place
# of average_optical_depth.cpp and planck_function.cpp and provides radiative_transfer_lw<false>; the stock
# radiative_transfer_lw.cpp is compiled with that one name changed so that its other functions stay.  Everything
# else - the mains, the file layer, every other unit - is the object code of the stock build above.
REPO="$(cd "$HERE/.." && pwd)"
if [ -f "$REPO/ecckd_b200/lib/libecckd_b200.so" ]; then
  g++ $FLAGS -I"$REPO/include" -c "$REPO/integration/ecckd_gpu_bindings.cpp" -o "$OUT/tools_obj/gpu_bindings.o" &
  g++ $FLAGS -fopenmp -Dradiative_transfer_lw=radiative_transfer_lw_stock -c "$REF/src/ecckd/radiative_transfer_lw.cpp" \
      -o "$OUT/tools_obj/radiative_transfer_lw_renamed.o" &
  gcc -std=c99 -Wall -Werror -fsyntax-only -I"$REPO/include" -x c "$REPO/include/ecckd_b200.h" &   # the header is C
  wait
  GPU_OBJS=""
  for o in $OBJS; do
    case "$o" in
      */average_optical_depth.o|*/planck_function.o) ;;
      */radiative_transfer_lw.o) GPU_OBJS="$GPU_OBJS $OUT/tools_obj/radiative_transfer_lw_renamed.o" ;;
      *) GPU_OBJS="$GPU_OBJS $o" ;;
    esac
  done
  g++ -shared -fopenmp -o "$OUT/libecckd_ref_tools_gpu.so" "$OUT/tools_obj/ref_tools_glue.o" "$OUT/tools_obj/ref_tools_fgp_glue.o" \
      "$OUT/tools_obj/mock_files.o" "$OUT/tools_obj/gpu_bindings.o" $GPU_OBJS \
      -L"$REPO/ecckd_b200/lib" -lecckd_b200 -Wl,-rpath,'$ORIGIN/../../ecckd_b200/lib'
  echo "built $OUT/libecckd_ref_tools_gpu.so"
else
  echo "libecckd_b200.so not built yet: skipping libecckd_ref_tools_gpu.so"
fi
rm -rf "$OUT/tools_obj"
