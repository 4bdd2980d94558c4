#!/bin/bash
# round 2, GPU call 29: reduce2_kernel without the optical-depth clamp on the fast path
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p29
LIB=ecckd_b200/lib/libecckd_b200.so
cp $LIB /tmp/lib_orig.so
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-flux --no-full"
for v in nc base nc; do
  cp ecckd_b200/lib/variants/lib_$v.so $LIB
  timeout 600 python -m pytest tests/test_reduce2_gpu.py tests/test_reduce_gpu.py -m gpu -x -q --timeout 300 > ${O}_tests_$v.log 2>&1; echo "$v tests rc=$?"; tail -1 ${O}_tests_$v.log
  timeout 600 $B > ${O}_$v.json 2> ${O}_$v.err
  python tools/r02_line.py ${O}_$v.json $v
done
cp /tmp/lib_orig.so $LIB
