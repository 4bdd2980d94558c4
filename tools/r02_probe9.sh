#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_flux_gpu.py -x -q -m gpu > gpurun_out/r02p9_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02p9_tests.log
tail -5 gpurun_out/r02p9_tests.log
timeout 300 python tools/run_ckd_sweep.py synthetic > gpurun_out/r02p9_runckd.log 2>&1
timeout 300 python tools/run_ckd_sweep.py coherent >> gpurun_out/r02p9_runckd.log 2>&1
cat gpurun_out/r02p9_runckd.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:total_od_planck -c 1 -o gpurun_out/r02p9_tod -f python tools/run_ckd_sweep.py synthetic one > gpurun_out/r02p9_ncu.log 2>&1
tail -3 gpurun_out/r02p9_ncu.log
