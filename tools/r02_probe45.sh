#!/bin/bash
# round 2, GPU call 45: source-level profile of plan_colour_kernel
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p45
timeout 900 ncu --set full --clock-control none --import-source on -k regex:plan_colour_kernel -s 2 -c 1 -o ${O}_colour -f \
  python tools/plan_hash.py > ${O}_ncu.log 2>&1
tail -2 ${O}_ncu.log
