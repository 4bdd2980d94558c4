#!/bin/bash
# round 2, GPU call 39: look-up-table adjoint with a warp-wide record test
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p39
timeout 900 python -m pytest tests/test_flux_gpu.py tests/test_cost_gpu.py -m gpu -x -q --timeout 600 -k "adjoint or gradient" > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -3 ${O}_tests.log
timeout 900 python tools/bench_extra.py > ${O}_bench_extra.json 2> ${O}_bench_extra.err; tail -3 ${O}_bench_extra.err; python tools/r02_show.py ${O}_bench_extra.json
