#!/bin/bash
# round 2, GPU call 59: sort key with the group depth chosen from the number of layers; ranking + scale tests
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p59
timeout 200 python tools/sort_key_profile.py 23 2>&1 | tail -1 | tee ${O}_sk.log
timeout 600 python -m pytest tests/test_ranking_gpu.py tests/test_scale_gpu.py tests/test_partition_search_gpu.py -m gpu -x -q 2>&1 | tail -2
