"""The host's aggregate pinned-memory -> device ceiling, all ranks copying at once (torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/h2d_ceiling.py
Every rank copies blocks of the size of one float32 [53][2^23] spectrum block (1.78 GB, what one end-to-end step of
bench.py uploads) from two alternating pinned buffers, timed on the device between barriers, first with the pages
wherever the process happened to touch them, then (``bound``) after pinning the process to the cores of the GPU's NUMA
node before the buffers are allocated and first touched.  Prints per-rank and aggregate GB/s: the ceiling the
end-to-end line of an N-GPU run can reach."""
import json
import os

import torch
import torch.distributed as dist


def numa_cpus_of_gpu(index):
    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read())
        if node < 0:
            return None, node
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return cpus & os.sched_getaffinity(0), node
    except Exception:  # no sysfs entry in this container
        return None, -1


def measure(dev, nbytes, reps):
    host = [torch.empty(nbytes // 4, dtype=torch.float32).pin_memory() for _ in range(2)]
    for h in host:
        h.fill_(1.0)  # first touch
    d = torch.empty(nbytes // 4, dtype=torch.float32, device=dev)
    for h in host:
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        d.copy_(host[i & 1], non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return nbytes * reps / (ms.item() * 1e-3) / 1e9  # GB/s per rank at the slowest rank's time


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29541")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    nbytes = 53 * (1 << 23) * 4
    out = {"world": world, "block_bytes": nbytes, "host_cpus": len(os.sched_getaffinity(0))}
    out["unbound_GBps_per_rank"] = measure(dev, nbytes, 8)
    cpus, node = numa_cpus_of_gpu(local)
    nodes = [None] * world
    dist.all_gather_object(nodes, node)
    out["numa_node_of_gpu"] = nodes
    if cpus:
        os.sched_setaffinity(0, cpus)
    out["bound_GBps_per_rank"] = measure(dev, nbytes, 8)
    for k in ("unbound", "bound"):
        out[k + "_GBps_aggregate"] = out[k + "_GBps_per_rank"] * world
    if rank == 0:
        print(json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
