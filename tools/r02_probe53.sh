#!/bin/bash
# round 2, GPU call 53 (8 GPUs): the driver's own 8-GPU command (weak, end-to-end legs included), the strong line and
# the bits of the outputs, on the final code
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p53
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
timeout 600 $TR bench.py --gpus 8 --steps 10 --warmup 3 > ${O}_n8_weak.json 2> ${O}_n8_weak.err; python tools/r02_line.py ${O}_n8_weak.json n8_weak
timeout 400 $TR bench.py --gpus 8 --steps 10 --warmup 3 --scaling strong > ${O}_n8_strong.json 2> ${O}_n8_strong.err; python tools/r02_line.py ${O}_n8_strong.json n8_strong
timeout 300 $TR tools/multi_gpu_bits.py > ${O}_bits_n8.log 2>&1; tail -2 ${O}_bits_n8.log
