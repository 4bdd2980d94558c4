#!/bin/bash
# round 2, GPU call 56: full GPU suite + smoke on the final tree
set -x
mkdir -p gpurun_out
O=gpurun_out/r02h
timeout 1500 python -m pytest tests -x -q -m gpu > ${O}_gpu_tests.log 2>&1; echo "tests rc=$?" >> ${O}_gpu_tests.log; tail -4 ${O}_gpu_tests.log
timeout 600 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' >> ${O}_gpu_tests.log 2>&1; tail -1 ${O}_gpu_tests.log
