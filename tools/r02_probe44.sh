#!/bin/bash
# round 2, GPU call 44: colouring kernel writes the cells slot-parallel; window kernel off the memory latencies
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p44
timeout 300 python tools/plan_hash.py > ${O}_hash.log 2>&1; cat ${O}_hash.log | tail -4
timeout 900 python -m pytest tests/test_reduce2_gpu.py tests/test_reduce_gpu.py tests/test_lut_builder_gpu.py -m gpu -x -q > ${O}_tests.log 2>&1; tail -3 ${O}_tests.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:plan_ --csv --log-file ${O}_plan_launches.csv python tools/plan_hash.py > ${O}_ncu.log 2>&1
tail -8 ${O}_plan_launches.csv | cut -d, -f5,15-
