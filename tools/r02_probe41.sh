#!/bin/bash
# round 2, GPU call 41: source-level profile of plan_fill_balanced_kernel
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p41
timeout 900 ncu --set full --clock-control none --import-source on -k regex:plan_fill_balanced_kernel -s 2 -c 1 -o ${O}_plan -f \
  python tools/plan_hash.py > ${O}_ncu.log 2>&1
tail -2 ${O}_ncu.log
