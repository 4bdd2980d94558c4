#!/bin/bash
# round 2, GPU call 33: interval flux kernel with transposed per-level sums
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p33
timeout 900 python -m pytest tests/test_cost_gpu.py tests/test_partition_search_gpu.py tests/test_flux_gpu.py tests/test_golden_reference.py -m gpu -x -q --timeout 600 > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -3 ${O}_tests.log
timeout 600 python tools/bench_extra.py > ${O}_bench_extra.json 2> ${O}_bench_extra.err; python -c "
import json; d=json.load(open('${O}_bench_extra.json')); print(d['interval_error_lw'])"
