#!/bin/bash
# round 2, GPU call 25: weight-sum duty on a producer scheduler, two-instruction table index
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p25
timeout 900 python -m pytest tests/test_reduce2_gpu.py tests/test_reduce_gpu.py tests/test_lut_builder_gpu.py tests/test_segments_gpu.py -m gpu -x -q --timeout 600 > ${O}_tests.log 2>&1; echo "tests rc=$?"; tail -3 ${O}_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-flux --no-cpu"
timeout 600 $B > ${O}_base.json 2> ${O}_base.err; python tools/r02_line.py ${O}_base.json s1_moved
ECCKD_S1_LAST=1 timeout 600 $B > ${O}_s1last.json 2> ${O}_s1last.err; python tools/r02_line.py ${O}_s1last.json s1_last
timeout 600 $B > ${O}_base2.json 2> ${O}_base2.err; python tools/r02_line.py ${O}_base2.json s1_moved_again
