#!/bin/bash
# round 2, GPU call 43: launch list of the plan build
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p43
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:plan_ --csv --log-file ${O}_plan_launches.csv python tools/plan_hash.py > ${O}_ncu.log 2>&1
tail -2 ${O}_ncu.log
