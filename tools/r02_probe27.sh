#!/bin/bash
# round 2, GPU call 27: the default bench with the full-build leg
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p27
timeout 900 python bench.py > ${O}_bench_line.json 2> ${O}_bench_line.err; echo "bench rc=$?"; python tools/r02_line.py ${O}_bench_line.json bench
python -c "
import json
d=json.loads(open('${O}_bench_line.json').read().strip().splitlines()[-1]); print(json.dumps(d.get('full_build'), indent=1)); print(d['flux_solver']['ms_per_step'], d['flux_solver']['roofline']['frac'])"
tail -5 ${O}_bench_line.err
