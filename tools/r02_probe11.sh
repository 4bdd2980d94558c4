#!/bin/bash
# round 2, GPU call 11 (2 GPUs): new tests, NCCL exchange bits for N=1/2, bench weak N=2 and strong N=1/2
set -x
mkdir -p gpurun_out
L=gpurun_out/r02p11
timeout 900 python -m pytest tests/test_segments_gpu.py tests/test_flux_gpu.py tests/test_lut_builder_gpu.py -x -q -m gpu > $L.tests.log 2>&1; echo "tests rc=$?" >> $L.tests.log
tail -5 $L.tests.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for m in transmission logarithmic; do
  timeout 300 python tools/multi_gpu_bits.py $m >> $L.bits.log 2>&1
  timeout 300 $TR --nproc-per-node 2 --master-port 29511 tools/multi_gpu_bits.py $m >> $L.bits.log 2>&1
done
grep "sha256\|Error\|error" $L.bits.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-flux > $L.weak1.json 2> $L.weak1.err; echo rc=$?
timeout 600 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-flux --no-cpu > $L.weak2.json 2> $L.weak2.err; echo rc=$?
timeout 600 python bench.py --scaling strong --steps 10 --warmup 3 --no-flux --no-e2e --no-cpu > $L.strong1.json 2> $L.strong1.err; echo rc=$?
timeout 600 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --scaling strong --steps 10 --warmup 3 --no-flux --no-e2e --no-cpu > $L.strong2.json 2> $L.strong2.err; echo rc=$?
for f in weak1 weak2 strong1 strong2; do echo "== $f"; python - <<PY
import json
try:
    d = json.loads(open("$L.$f.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling")}, d["roofline"]["frac"], d["roofline"]["kernel_ms"], (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("no json", e)
PY
tail -3 $L.$f.err; done
