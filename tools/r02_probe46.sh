#!/bin/bash
# round 2, GPU call 46: what the Planck look-up table leg launches
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p46
timeout 300 python tools/planck_lut_probe.py > ${O}_plain.log 2>&1; tail -4 ${O}_plain.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches.csv python tools/planck_lut_probe.py > ${O}_ncu.log 2>&1
tail -2 ${O}_ncu.log
