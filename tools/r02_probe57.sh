#!/bin/bash
# round 2, GPU call 57: longwave sort key with 3 / 6 / 9 / 18 layers per group (loads in flight, independent chains)
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p57
for g in 6 3 9 18; do echo "SK_G=$g"; ECCKD_SK_G=$g timeout 200 python tools/sort_key_profile.py 23 2>&1 | tail -1; done | tee ${O}_skg.log
ECCKD_SK_G=9 timeout 300 python -m pytest tests/test_ranking_gpu.py -m gpu -x -q 2>&1 | tail -2
