set -x
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 --durations=8 > $O/r02f_tests.log 2>&1; echo "tests rc=$?" >> $O/r02f_tests.log
tail -25 $O/r02f_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-flux"
run() { name=$1; shift; timeout 600 env "$@" $B $EXTRA > $O/r02f_$name.json 2> $O/r02f_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/r02f_$name.json").read().strip().splitlines()[-1])
    print("$name: %.1f G/s kernel_ms %.3f frac %.3f" % (d['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac']))
except Exception as e:
    print("$name: FAILED", e)
PY
}
EXTRA="" run g2_nsw X=1
EXTRA="--gmap bands" run g2_nsw_bands X=1
EXTRA="--method linear" run g2_nsw_linear X=1
