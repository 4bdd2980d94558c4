#!/bin/bash
# round 2, GPU call 20: floating row in reduce2_kernel, transposed broadband sums in the staged longwave sweep
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p20
timeout 900 python -m pytest tests/test_reduce2_gpu.py tests/test_reduce_gpu.py tests/test_flux_gpu.py tests/test_scale_gpu.py tests/test_lut_builder_gpu.py tests/test_segments_gpu.py -m gpu -x -q --timeout 600 > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -3 ${O}_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-flux"
timeout 600 $B > ${O}_float.json 2> ${O}_float.err; python tools/r02_line.py ${O}_float.json float_row
ECCKD_NO_FLOAT_ROW=1 timeout 600 $B > ${O}_nofloat.json 2> ${O}_nofloat.err; python tools/r02_line.py ${O}_nofloat.json no_float_row
timeout 600 $B --method linear > ${O}_linear.json 2> ${O}_linear.err; python tools/r02_line.py ${O}_linear.json linear
timeout 600 $B --nspec 20 > ${O}_n20.json 2> ${O}_n20.err; python tools/r02_line.py ${O}_n20.json nspec20
timeout 300 python tools/run_ckd_sweep.py synthetic > ${O}_run_ckd.log 2>&1; tail -2 ${O}_run_ckd.log
