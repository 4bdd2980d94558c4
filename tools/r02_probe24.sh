#!/bin/bash
# round 2, GPU call 24: table-based Planck weights in the producers of reduce2_kernel; OD kernel variant 3
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p24
timeout 900 python -m pytest tests/test_reduce2_gpu.py tests/test_reduce_gpu.py tests/test_scale_gpu.py tests/test_lut_builder_gpu.py tests/test_segments_gpu.py tests/test_golden_reference.py -m gpu -x -q --timeout 600 > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -3 ${O}_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-flux"
timeout 600 $B > ${O}_base.json 2> ${O}_base.err; python tools/r02_line.py ${O}_base.json planck_tab
python -c "
import json
d=json.loads(open('${O}_base.json').read().strip().splitlines()[-1]); print(d['validated'])"
timeout 600 $B --no-cpu --weights matrix > ${O}_matrix.json 2> ${O}_matrix.err; python tools/r02_line.py ${O}_matrix.json matrix_weights
timeout 300 python tools/run_ckd_sweep.py synthetic > ${O}_run_ckd.log 2>&1; tail -4 ${O}_run_ckd.log
