#!/bin/bash
# round 2, GPU call 31: total_od_planck3_kernel with the L1 prefetch of the column after next
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p31
timeout 900 python -m pytest tests/test_flux_gpu.py -m gpu -x -q --timeout 600 > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -2 ${O}_tests.log
timeout 300 python tools/run_ckd_sweep.py synthetic > ${O}_run_ckd.log 2>&1; tail -4 ${O}_run_ckd.log
timeout 300 python tools/run_ckd_sweep.py coherent > ${O}_run_ckd_coh.log 2>&1; tail -2 ${O}_run_ckd_coh.log
