#!/bin/bash
# round 2, GPU call 28: lw_spectra / scale_lut entry points, floating-row test
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p28
timeout 900 python -m pytest tests/test_spectra_tools_gpu.py tests/test_reduce2_gpu.py tests/test_stage1_misc_gpu.py -m gpu -x -q --timeout 600 > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -15 ${O}_tests.log
