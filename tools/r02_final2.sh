#!/bin/bash
# round 2, second evidence run (after the plan colouring kernel, planck_lut_kernel and the vectorised merge; the reduce
# and run_ckd kernels are those of tools/r02_final.sh): full GPU suite + smoke, default bench line, launch lists,
# ncu --set full of planck_lut_kernel.  Outputs under gpurun_out/r02g_*.
set -x
mkdir -p gpurun_out
O=gpurun_out/r02g
timeout 2400 python -m pytest tests -x -q -m gpu > ${O}_gpu_tests.log 2>&1; echo "tests rc=$?" >> ${O}_gpu_tests.log; tail -4 ${O}_gpu_tests.log
timeout 600 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' >> ${O}_gpu_tests.log 2>&1; tail -1 ${O}_gpu_tests.log
timeout 900 python bench.py > ${O}_bench_line.json 2> ${O}_bench_line.err; echo "bench rc=$?"; python tools/r02_line.py ${O}_bench_line.json bench
K='regex:reduce2_kernel|reduce_kernel|combine_kernel|finalize_kernel|plan_|planck_|merge_spectra|gfrac|hist_max|fix128|compact_|total_od|rt_lw|rt_sw|run_ckd|sum_g|widen_f32|stage_copy|g_point_differs|exclusive_scan|segment_combine|upload'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 900 --csv --log-file ${O}_bench_launch_list.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu > ${O}_ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:plan_ --csv --log-file ${O}_plan_launches.csv python tools/plan_hash.py > ${O}_plan_hash.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:planck_ --csv --log-file ${O}_planck_lut_launches.csv python tools/planck_lut_probe.py > ${O}_planck_probe_ncu.log 2>&1
timeout 300 python tools/planck_lut_probe.py > ${O}_planck_probe.log 2>&1; tail -8 ${O}_planck_probe.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:planck_lut_kernel -s 1 -c 1 -o ${O}_plut -f \
  python tools/planck_lut_probe.py > ${O}_ncu_plut.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:plan_colour_kernel -s 2 -c 1 -o ${O}_colour -f \
  python tools/plan_hash.py > ${O}_ncu_colour.log 2>&1
ls -la gpurun_out | grep r02g
