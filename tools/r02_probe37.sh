#!/bin/bash
# round 2, GPU call 37: the full GPU suite and smoke() on the final tree
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p37
timeout 2400 python -m pytest tests -x -q -m gpu > ${O}_gpu_tests.log 2>&1; echo "tests rc=$?" >> ${O}_gpu_tests.log; tail -5 ${O}_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 ${O}_smoke.log
