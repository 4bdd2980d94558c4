set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_reduce2_gpu.py tests/test_reduce_gpu.py -m gpu -x -q --timeout 300 > $O/r02c_tests.log 2>&1; echo "tests rc=$?" >> $O/r02c_tests.log
tail -8 $O/r02c_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-flux"
run() { name=$1; shift; timeout 600 env "$@" $B $EXTRA > $O/r02c_$name.json 2> $O/r02c_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/r02c_$name.json").read().strip().splitlines()[-1])
    print("$name: %.1f G/s kernel_ms %.3f frac %.3f" % (d['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac']))
except Exception as e:
    print("$name: FAILED", e)
PY
}
EXTRA="" run g2_bal ECCKD_REDUCE_GEN=2
EXTRA="" run g3_bal X=1
EXTRA="" run g2_bal480 ECCKD_REDUCE_GEN=2 ECCKD_CH=480
EXTRA="--method linear" run g2_bal_linear ECCKD_REDUCE_GEN=2
EXTRA="--method linear" run g3_bal_linear X=1
EXTRA="--gmap random" run g2_bal_random ECCKD_REDUCE_GEN=2
timeout 900 env ECCKD_REDUCE_GEN=2 ncu --set full --clock-control none --import-source on -k regex:reduce2_kernel -s 3 -c 1 -o $O/r02c_reduce2 -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-flux > $O/r02c_ncu_full.log 2>&1
