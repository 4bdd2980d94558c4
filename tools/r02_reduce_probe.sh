# Round-2 probe of the spectral reduction on the GPU box (run from the repo root under gpurun):
# parity tests, both kernel generations on the bench workload, then the ncu captures.
set -x
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 240 > $O/r02_tests.log 2>&1; echo "tests rc=$?" >> $O/r02_tests.log
tail -5 $O/r02_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-flux"
run() { name=$1; shift; timeout 600 env "$@" $B $EXTRA > $O/r02_$name.json 2> $O/r02_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/r02_$name.json").read().strip().splitlines()[-1])
    print("$name: %.1f G/s kernel_ms %.3f frac %.3f" % (d['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac']))
except Exception as e:
    print("$name: FAILED", e)
PY
}
EXTRA="" run v3 X=1
EXTRA="" run v2 ECCKD_REDUCE_GEN=2
EXTRA="--method linear" run v3_linear X=1
EXTRA="--gmap random" run v3_random X=1
EXTRA="--gmap bands" run v3_bands X=1
EXTRA="--weights matrix" run v3_matrix X=1
# ncu: launch list of the bench command, then one full capture of the reduce kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'reduce|combine|finalize|plan_' --csv \
  --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-flux > $O/r02_ncu_ll.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reduce3_kernel -s 3 -c 1 -o $O/r02_reduce3 -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-flux > $O/r02_ncu_full.log 2>&1
ls -la $O | tail -20
