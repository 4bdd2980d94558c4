#!/bin/bash
# round 2, GPU call 58: source-level profile of sort_key_lw_kernel<18> at 54 x 2^22
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p58
ECCKD_SK_G=18 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sort_key_lw_kernel -s 1 -c 1 -o ${O}_sk -f \
  python tools/sort_key_profile.py 22 > ${O}_ncu.log 2>&1
tail -2 ${O}_ncu.log
