#!/bin/bash
# round 2, GPU call 38: bench_extra with the gradient legs
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p38
timeout 900 python tools/bench_extra.py > ${O}_bench_extra.json 2> ${O}_bench_extra.err; tail -3 ${O}_bench_extra.err; python -c "
import json; d=json.load(open('${O}_bench_extra.json')); print({k:d[k] for k in d if 'grad' in k}); print(d['forward_cost_lw_batched'])"
