#!/bin/bash
# round 2, GPU call 52: colouring that takes a colour free at both ends first (different plans: reduce kernel time? conflicts?)
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p52
timeout 300 python tools/plan_hash.py > ${O}_hash.log 2>&1; cat ${O}_hash.log | tail -4
timeout 900 python -m pytest tests/test_reduce2_gpu.py tests/test_reduce_gpu.py tests/test_lut_builder_gpu.py -m gpu -x -q > ${O}_tests.log 2>&1; tail -3 ${O}_tests.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-flux --no-e2e > ${O}_bench.log 2>&1; tail -1 ${O}_bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['full_build']['ms'], d['full_build']['phases_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:'plan_colour|reduce2_kernel' -c 6 --csv --log-file ${O}_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-flux --no-e2e --no-full > ${O}_ncu.log 2>&1
grep -v "^==" ${O}_launches.csv | cut -d, -f5,13- | cut -c1-40,100- | tail -12
