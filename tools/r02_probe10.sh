#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/dbg_tod.py > gpurun_out/r02p10_dbg.log 2>&1
cat gpurun_out/r02p10_dbg.log
timeout 300 python tools/run_ckd_sweep.py synthetic > gpurun_out/r02p10_runckd.log 2>&1
cat gpurun_out/r02p10_runckd.log
