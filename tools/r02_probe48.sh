#!/bin/bash
# round 2, GPU call 48: source-level profile of planck_lut_kernel
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p48
timeout 900 ncu --set full --clock-control none --import-source on -k regex:planck_lut_kernel -s 1 -c 1 -o ${O}_plut -f \
  python tools/planck_lut_probe.py > ${O}_ncu.log 2>&1
tail -2 ${O}_ncu.log
