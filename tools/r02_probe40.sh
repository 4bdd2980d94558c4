#!/bin/bash
# round 2, GPU call 40: balanced plan kernel with bit-mask colour search: same plans, build time
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p40
LIB=ecckd_b200/lib/libecckd_b200.so
cp $LIB /tmp/lib_new.so
timeout 600 python tools/plan_hash.py > ${O}_new.log 2>&1; cat ${O}_new.log | tail -3
cp ecckd_b200/lib/variants/lib_before_plan.so $LIB
timeout 600 python tools/plan_hash.py > ${O}_old.log 2>&1; cat ${O}_old.log | tail -3
cp /tmp/lib_new.so $LIB
timeout 900 python -m pytest tests/test_reduce2_gpu.py tests/test_reduce_gpu.py tests/test_lut_builder_gpu.py -m gpu -x -q --timeout 600 > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -2 ${O}_tests.log
