#!/bin/bash
# round 2, GPU call 23: total_od_planck2_kernel two columns per trip
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p23
timeout 900 python -m pytest tests/test_flux_gpu.py -m gpu -x -q --timeout 600 > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -3 ${O}_tests.log
timeout 300 python tools/run_ckd_sweep.py synthetic > ${O}_run_ckd.log 2>&1; tail -3 ${O}_run_ckd.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'total_od_planck2_kernel' -s 1 -c 1 -o ${O}_tod2 -f \
  python tools/run_ckd_sweep.py synthetic one > ${O}_ncu_tod2.log 2>&1
tail -2 ${O}_ncu_tod2.log
