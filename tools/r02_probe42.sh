#!/bin/bash
# round 2, GPU call 42: plan build split into fill + one-task-per-warp colouring: plans bit-identical?  build time?
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p42
timeout 300 python tools/plan_hash.py > ${O}_hash.log 2>&1; cat ${O}_hash.log | tail -4
timeout 900 python -m pytest tests/test_reduce2_gpu.py tests/test_reduce_gpu.py tests/test_lut_builder_gpu.py -m gpu -x -q > ${O}_tests.log 2>&1; tail -3 ${O}_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-flux > ${O}_bench.log 2>&1; tail -1 ${O}_bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d.get('full_build'))"
