#!/bin/bash
# round 2, GPU call 21: total_od_planck2_kernel (asynchronous staging), rotating pipeline variant of reduce2_kernel
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p21
timeout 900 python -m pytest tests/test_flux_gpu.py tests/test_scale_gpu.py tests/test_cxx_consumer_gpu.py tests/test_golden_reference.py -m gpu -x -q --timeout 600 > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -5 ${O}_tests.log
timeout 300 python tools/run_ckd_sweep.py synthetic > ${O}_run_ckd.log 2>&1; tail -3 ${O}_run_ckd.log
timeout 300 python tools/run_ckd_sweep.py coherent > ${O}_run_ckd_coh.log 2>&1; tail -3 ${O}_run_ckd_coh.log
LIB=ecckd_b200/lib/libecckd_b200.so
cp $LIB /tmp/lib_orig.so
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-flux"
timeout 600 $B > ${O}_base.json 2> ${O}_base.err; python tools/r02_line.py ${O}_base.json base
cp ecckd_b200/lib/variants/lib_rot.so $LIB
timeout 600 python -m pytest tests/test_reduce2_gpu.py -m gpu -x -q --timeout 300 > ${O}_tests_rot.log 2>&1; echo "rot tests rc=$?"
timeout 600 $B > ${O}_rot.json 2> ${O}_rot.err; python tools/r02_line.py ${O}_rot.json rot
cp /tmp/lib_orig.so $LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'total_od_planck2_kernel' -s 2 -c 1 -o ${O}_tod2 -f \
  python tools/run_ckd_sweep.py synthetic one > ${O}_ncu_tod2.log 2>&1
tail -2 ${O}_ncu_tod2.log
