#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cxx_consumer_gpu.py -x -q -m gpu > gpurun_out/r02p12_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02p12_tests.log
tail -30 gpurun_out/r02p12_tests.log
