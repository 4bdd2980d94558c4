#!/bin/bash
# round 2, GPU call 49: full GPU suite + bench after the plan / Planck look-up table kernels
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p49
timeout 1500 python -m pytest tests -m gpu -x -q > ${O}_tests.log 2>&1; tail -3 ${O}_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-flux > ${O}_bench.log 2>&1; tail -1 ${O}_bench.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d.get('full_build'))"
