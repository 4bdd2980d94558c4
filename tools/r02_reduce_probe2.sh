# Round-2 probe: balanced plans (run from the repo root under gpurun)
set -x
O=gpurun_out
timeout 900 python -m pytest tests/test_reduce2_gpu.py tests/test_reduce_gpu.py tests/test_trans_core_gpu.py -m gpu -x -q --timeout 300 > $O/r02b_tests.log 2>&1; echo "tests rc=$?" >> $O/r02b_tests.log
tail -15 $O/r02b_tests.log
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-flux"
run() { name=$1; shift; timeout 600 env "$@" $B $EXTRA > $O/r02b_$name.json 2> $O/r02b_$name.err; python - <<PY
import json
try:
    d=json.loads(open("$O/r02b_$name.json").read().strip().splitlines()[-1])
    print("$name: %.1f G/s kernel_ms %.3f frac %.3f" % (d['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac']))
except Exception as e:
    print("$name: FAILED", e)
PY
}
EXTRA="" run g2_bal ECCKD_REDUCE_GEN=2
EXTRA="" run g2_unbal ECCKD_REDUCE_GEN=2 ECCKD_PLAN_BALANCE=0
EXTRA="" run g3_bal X=1
EXTRA="--method linear" run g2_bal_linear ECCKD_REDUCE_GEN=2
EXTRA="--gmap random" run g2_bal_random ECCKD_REDUCE_GEN=2
EXTRA="--gmap random" run g3_bal_random X=1
EXTRA="--gmap bands" run g2_bal_bands ECCKD_REDUCE_GEN=2
EXTRA="--weights matrix" run g2_bal_matrix ECCKD_REDUCE_GEN=2
timeout 900 env ECCKD_REDUCE_GEN=2 ncu --set full --clock-control none --import-source on -k regex:reduce2_kernel -s 3 -c 1 -o $O/r02b_reduce2 -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-flux > $O/r02b_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'plan_' --csv \
  --log-file $O/r02b_plan_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-flux > $O/r02b_ncu_ll.log 2>&1
grep -c reduce $O/r02b_plan_launches.csv; tail -12 $O/r02b_plan_launches.csv | cut -c1-200
