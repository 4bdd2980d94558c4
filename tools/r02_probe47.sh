#!/bin/bash
# round 2, GPU call 47: Planck look-up table with the temperatures across the lanes
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p47
timeout 300 python tools/planck_lut_probe.py > ${O}_plain.log 2>&1; tail -8 ${O}_plain.log
timeout 600 python -m pytest tests/test_stage1_misc_gpu.py tests/test_lut_builder_gpu.py -m gpu -x -q > ${O}_tests.log 2>&1; tail -3 ${O}_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:planck_ --csv --log-file ${O}_launches.csv python tools/planck_lut_probe.py > ${O}_ncu.log 2>&1
grep planck_ ${O}_launches.csv | cut -d, -f5,15- | tail -6
