#!/bin/bash
# round 2, GPU call 55: background (prior) term against the oracle
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p55
timeout 600 python -m pytest tests/test_cost_gpu.py -m gpu -x -q > ${O}_tests.log 2>&1; tail -5 ${O}_tests.log
