#!/bin/bash
# round 2: the evidence run - full GPU suite, default bench line, reference arm, launch list, ncu --set full of the
# headline kernel and of the run_ckd kernels, the rows either side of the path.  Outputs under gpurun_out/r02f_*.
set -x
mkdir -p gpurun_out
O=gpurun_out/r02f
timeout 2400 python -m pytest tests -x -q -m gpu > ${O}_gpu_tests.log 2>&1; echo "tests rc=$?" >> ${O}_gpu_tests.log; tail -4 ${O}_gpu_tests.log
timeout 900 python bench.py > ${O}_bench_line.json 2> ${O}_bench_line.err; echo "bench rc=$?"; python tools/r02_line.py ${O}_bench_line.json bench
timeout 900 python bench.py --impl reference > ${O}_bench_line_reference.json 2> ${O}_bench_line_reference.err; echo "reference rc=$?"
K='regex:reduce2_kernel|reduce_kernel|combine_kernel|finalize_kernel|plan_|total_od|rt_lw|rt_sw|run_ckd|sum_g|widen_f32|stage_copy|g_point_differs|exclusive_scan|segment_combine|upload'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 600 --csv --log-file ${O}_bench_launch_list.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > ${O}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reduce2_kernel -s 3 -c 1 -o ${O}_reduce2 -f \
  python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-flux > ${O}_ncu_reduce2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'total_od_planck3_kernel|rt_lw_staged_kernel' -s 2 -c 2 -o ${O}_run_ckd -f \
  python tools/run_ckd_sweep.py synthetic one > ${O}_ncu_run_ckd.log 2>&1
timeout 600 python tools/bench_extra.py > ${O}_bench_extra.json 2> ${O}_bench_extra.err; tail -c 600 ${O}_bench_extra.json
ls -la gpurun_out | grep r02f
