#!/bin/bash
# round 2, GPU call 15: reduce2_kernel with 11 consumer warps (rows balanced over the schedulers) against 10
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p15
timeout 900 python -m pytest tests/test_reduce2_gpu.py tests/test_reduce_gpu.py tests/test_lut_builder_gpu.py -m gpu -x -q --timeout 300 > ${O}_tests.log 2>&1; echo "tests rc=$?" >> ${O}_tests.log
tail -5 ${O}_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-flux"
run() { name=$1; shift; timeout 600 env "$@" $B $EXTRA > ${O}_$name.json 2> ${O}_$name.err; python - <<PY
import json
try:
    d=json.loads(open("${O}_$name.json").read().strip().splitlines()[-1])
    print("$name: %.1f G/s kernel_ms %.3f frac %.3f valid %s" % (d['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac'], (d.get('validated') or {}).get('max_rel_err')))
except Exception as e:
    print("$name: FAILED", e)
PY
}
EXTRA="" run nw11 ECCKD_REDUCE_NW=11
EXTRA="" run nw10 ECCKD_REDUCE_NW=10
EXTRA="" run nw9 ECCKD_REDUCE_NW=9
EXTRA="--method linear" run nw11_linear ECCKD_REDUCE_NW=11
M=smsp__inst_executed.sum,smsp__inst_executed.max,smsp__inst_executed.min,smsp__inst_executed.avg,smsp__issue_active.max,smsp__issue_active.min,smsp__issue_active.avg,smsp__cycles_active.avg,sm__inst_executed_pipe_fp64.max,sm__inst_executed_pipe_fp64.avg
for nw in 11 10; do
timeout 600 env ECCKD_REDUCE_NW=$nw ncu --metrics $M --clock-control none -k regex:reduce2_kernel -s 3 -c 1 --csv --log-file ${O}_smsp_nw$nw.csv \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-flux --log2-nwav 22 > ${O}_smsp_nw$nw.log 2>&1
done
timeout 900 env ECCKD_REDUCE_NW=11 ncu --set full --clock-control none --import-source on -k regex:reduce2_kernel -s 3 -c 1 -o ${O}_reduce2 -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-flux --log2-nwav 22 > ${O}_ncu_full.log 2>&1
ls -la gpurun_out | tail -20
