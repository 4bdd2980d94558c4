#!/bin/bash
# round 2, GPU call 51: plan build without the redundant last counting pass; Planck table test with holes
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p51
timeout 300 python tools/plan_hash.py > ${O}_hash.log 2>&1; cat ${O}_hash.log | tail -4
timeout 900 python -m pytest tests/test_reduce2_gpu.py tests/test_reduce_gpu.py tests/test_lut_builder_gpu.py tests/test_stage1_misc_gpu.py -m gpu -x -q > ${O}_tests.log 2>&1; tail -3 ${O}_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-flux --no-e2e > ${O}_bench.log 2>&1; tail -1 ${O}_bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['full_build']['ms'], d['full_build']['phases_ms'])"
