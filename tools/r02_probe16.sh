#!/bin/bash
# round 2, GPU call 16: source-level stall profile of reduce2_kernel (168 registers), run_ckd kernels under ncu
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p16
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-flux"
timeout 600 $B > ${O}_base.json 2> ${O}_base.err
python -c "
import json
d=json.loads(open('${O}_base.json').read().strip().splitlines()[-1])
print('base: %.1f G/s kernel_ms %.3f frac %.3f' % (d['value']/1e9, d['roofline']['kernel_ms'], d['roofline']['frac']))"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reduce2_kernel -s 3 -c 1 -o ${O}_reduce2 -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-flux --log2-nwav 22 > ${O}_ncu_full.log 2>&1
timeout 300 python tools/run_ckd_profile.py > ${O}_run_ckd.log 2>&1; tail -2 ${O}_run_ckd.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'total_od_planck_kernel|rt_lw_staged_kernel' -s 6 -c 2 -o ${O}_run_ckd -f \
  python tools/run_ckd_profile.py > ${O}_ncu_run_ckd.log 2>&1
ls -la gpurun_out | tail
