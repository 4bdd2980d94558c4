#!/bin/bash
# round 2, GPU call 30: ncu of total_od_planck3_kernel
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p30
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'total_od_planck3_kernel' -s 1 -c 1 -o ${O}_tod3 -f \
  python tools/run_ckd_sweep.py synthetic one > ${O}_ncu_tod3.log 2>&1
tail -2 ${O}_ncu_tod3.log
