#!/bin/bash
# round 2, GPU call 26 (2 GPUs): weak and strong scaling lines, bitwise N-independence
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p26
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > ${O}_n2_weak.json 2> ${O}_n2_weak.err; python tools/r02_line.py ${O}_n2_weak.json n2_weak
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong > ${O}_n2_strong.json 2> ${O}_n2_strong.err; python tools/r02_line.py ${O}_n2_strong.json n2_strong
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --scaling strong --no-e2e --no-cpu --no-flux > ${O}_n1_strong.json 2> ${O}_n1_strong.err; python tools/r02_line.py ${O}_n1_strong.json n1_strong
timeout 600 python tools/multi_gpu_bits.py > ${O}_bits_n1.log 2>&1; tail -3 ${O}_bits_n1.log
timeout 600 $TR tools/multi_gpu_bits.py > ${O}_bits_n2.log 2>&1; tail -3 ${O}_bits_n2.log
