#!/bin/bash
# round 2, GPU call 8: ncu capture of total_od_planck_kernel; config D scale test
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scale_gpu.py tests/test_flux_gpu.py tests/test_lut_builder_gpu.py -x -q -m gpu > gpurun_out/r02p8_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02p8_tests.log
tail -5 gpurun_out/r02p8_tests.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:total_od_planck -c 1 -o gpurun_out/r02p8_tod -f python tools/run_ckd_sweep.py synthetic one > gpurun_out/r02p8_ncu.log 2>&1
tail -3 gpurun_out/r02p8_ncu.log
