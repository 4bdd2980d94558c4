#!/bin/bash
# round 2, GPU call 13: full gpu suite (incl. partition search + C++ consumer), bench both arms, launch list, ncu full of reduce2
set -x
mkdir -p gpurun_out
L=gpurun_out/r02p13
timeout 1800 python -m pytest tests -x -q -m gpu > $L.tests.log 2>&1; echo "tests rc=$?" >> $L.tests.log
tail -8 $L.tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > $L.bench.json 2> $L.bench.err; echo "bench rc=$?"
tail -c 3000 $L.bench.json; tail -5 $L.bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $L.ref.json 2> $L.ref.err; echo "ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $L.launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $L.ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:reduce2_kernel -s 2 -c 1 -o $L.reduce2 -f python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-flux > $L.ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 $L.ncu_full.log
