#!/bin/bash
# round 2, GPU call 14: full gpu suite
set -x
mkdir -p gpurun_out
L=gpurun_out/r02p14
timeout 1800 python -m pytest tests -x -q -m gpu > $L.tests.log 2>&1; echo "tests rc=$?" >> $L.tests.log
tail -15 $L.tests.log
