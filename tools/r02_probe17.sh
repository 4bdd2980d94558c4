#!/bin/bash
# round 2, GPU call 17: reduce2_kernel variants (table / range reduction / offset prefetch), run_ckd kernels under ncu
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p17
LIB=ecckd_b200/lib/libecckd_b200.so
cp $LIB /tmp/lib_orig.so
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-flux"
for v in v0 v1 v3 v4 v5 v6; do
  cp ecckd_b200/lib/variants/lib_$v.so $LIB
  timeout 600 python -m pytest tests/test_reduce2_gpu.py tests/test_trans_core_gpu.py -m gpu -x -q --timeout 300 > ${O}_tests_$v.log 2>&1; echo "$v tests rc=$?"; tail -1 ${O}_tests_$v.log
  timeout 600 $B > ${O}_$v.json 2> ${O}_$v.err
  python tools/r02_line.py ${O}_$v.json $v
done
cp ecckd_b200/lib/variants/lib_v0.so $LIB
timeout 600 python -m pytest tests/test_stage1_misc_gpu.py tests/test_lut_builder_gpu.py -m gpu -x -q --timeout 300 > ${O}_tests_misc.log 2>&1; echo "misc tests rc=$?"; tail -3 ${O}_tests_misc.log
timeout 300 python tools/run_ckd_sweep.py synthetic > ${O}_run_ckd.log 2>&1; tail -3 ${O}_run_ckd.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'total_od_planck_kernel|rt_lw_staged_kernel' -s 2 -c 2 -o ${O}_run_ckd -f \
  python tools/run_ckd_sweep.py synthetic one > ${O}_ncu_run_ckd.log 2>&1
tail -3 ${O}_ncu_run_ckd.log
cp /tmp/lib_orig.so $LIB
