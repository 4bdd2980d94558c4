# Timing experiments on the reduce kernel (run on the GPU box from the repo root): sh tools/variants.sh [log2_nwav]
# The ECCKD_DEBUG legs need a library built with ECCKD_TIMING_KNOBS=1 python -c 'import __graft_entry__ as g; g.build_library(True)'
N=${1:-22}
run() { echo "== DEBUG=$ECCKD_DEBUG $*"; timeout 300 python bench.py --steps 3 --warmup 3 --log2-nwav $N --no-e2e --no-cpu --no-flux "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.1f G/s  kernel_ms %.3f frac %.3f' % (d['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac']))"; }
run
run --weights matrix
run --gmap bands
run --gmap random
run --method linear
ECCKD_DEBUG=1 run
ECCKD_DEBUG=4 run
