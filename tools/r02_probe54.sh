#!/bin/bash
# round 2, GPU call 54 (8 GPUs): the host's aggregate pinned -> device ceiling at 1, 2, 4, 8 ranks
set -x
mkdir -p gpurun_out
O=gpurun_out/r02p54
for n in 1 2 4 8; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n tools/h2d_ceiling.py 2> ${O}_n$n.err | tail -1 | tee -a ${O}_h2d_ceiling.jsonl
done
nvidia-smi topo -m > ${O}_topo.txt 2>&1; head -14 ${O}_topo.txt
