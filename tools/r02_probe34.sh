#!/bin/bash
# round 2, GPU call 34: hand-written reverse mode of the longwave forward cost against finite differences
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p34
timeout 900 python -m pytest tests/test_cost_gpu.py -m gpu -x -q --timeout 600 -k "gradient or cost_function_ckd_lw" > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -25 ${O}_tests.log
