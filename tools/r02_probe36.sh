#!/bin/bash
# round 2, GPU call 36: gradient chain end to end
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p36
timeout 900 python -m pytest tests/test_cost_gpu.py -m gpu -x -q --timeout 600 -k "gradient" > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -25 ${O}_tests.log
