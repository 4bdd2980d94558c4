set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_reduce2_gpu.py tests/test_reduce_gpu.py -m gpu -x -q --timeout 300 > $O/r02e_tests.log 2>&1; echo "tests rc=$?" >> $O/r02e_tests.log
tail -8 $O/r02e_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-flux"
run() { name=$1; shift; timeout 600 env "$@" $B $EXTRA > $O/r02e_$name.json 2> $O/r02e_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/r02e_$name.json").read().strip().splitlines()[-1])
    print("$name: %.1f G/s kernel_ms %.3f frac %.3f" % (d['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac']))
except Exception as e:
    print("$name: FAILED", e)
PY
}
EXTRA="" run g2 ECCKD_REDUCE_GEN=2
EXTRA="--method linear" run g2_linear ECCKD_REDUCE_GEN=2
EXTRA="--gmap bands" run g2_bands ECCKD_REDUCE_GEN=2
timeout 900 env ECCKD_REDUCE_GEN=2 ncu --set full --clock-control none --import-source on -k regex:reduce2_kernel -s 3 -c 1 -o $O/r02e_reduce2 -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-flux > $O/r02e_ncu_full.log 2>&1
