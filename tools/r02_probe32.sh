#!/bin/bash
# round 2, GPU call 32 (8 GPUs): weak and strong scaling lines of the final kernels
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p32
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
timeout 900 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e > ${O}_n8_weak.json 2> ${O}_n8_weak.err; python tools/r02_line.py ${O}_n8_weak.json n8_weak
timeout 900 $TR bench.py --gpus 8 --steps 10 --warmup 3 --scaling strong > ${O}_n8_strong.json 2> ${O}_n8_strong.err; python tools/r02_line.py ${O}_n8_strong.json n8_strong
timeout 600 $TR tools/multi_gpu_bits.py > ${O}_bits_n8.log 2>&1; tail -2 ${O}_bits_n8.log
