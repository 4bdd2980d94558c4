#!/bin/bash
# round 2, GPU call 7: full gpu test suite, the run_ckd slice sweep, the bench (both arms)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02p7_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02p7_tests.log
tail -5 gpurun_out/r02p7_tests.log
timeout 300 python tools/run_ckd_sweep.py synthetic > gpurun_out/r02p7_runckd.log 2>&1
timeout 300 python tools/run_ckd_sweep.py coherent >> gpurun_out/r02p7_runckd.log 2>&1
cat gpurun_out/r02p7_runckd.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02p7_bench.json 2> gpurun_out/r02p7_bench.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/r02p7_bench.json; tail -20 gpurun_out/r02p7_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02p7_ref.json 2> gpurun_out/r02p7_ref.err; echo "ref rc=$?"
cat gpurun_out/r02p7_ref.json
