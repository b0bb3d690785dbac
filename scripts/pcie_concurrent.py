#!/usr/bin/env python
"""Concurrent pinned D2H ceiling of the node: every rank copies `mb` MB device->host at the same time (what bench.py's
e2e leg does at N GPUs).  Modes: default placement of the pinned buffer, and the process bound to the GPU's local
CPUs (sysfs local_cpulist) BEFORE the buffer is allocated (first touch puts the pages on that NUMA node).
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/pcie_concurrent.py"""
import os
import time

import torch
import torch.distributed as dist


def local_cpus(dev):
    try:
        bus = torch.cuda.get_device_properties(dev).pci_bus_id
        dom = torch.cuda.get_device_properties(dev).pci_domain_id
        devid = torch.cuda.get_device_properties(dev).pci_device_id
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0" % (dom, bus, devid)
        cpus = open(path + "/local_cpulist").read().strip()
        node = open(path + "/numa_node").read().strip()
        out = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            out.update(range(int(lo), int(hi or lo) + 1))
        return out, node, cpus
    except Exception as e:                                   # noqa: BLE001
        return None, "?", str(e)


def measure(dev, mb, reps, world):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([n / dt / 1e9], dtype=torch.float64, device=dev)
    lo = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    return float(t), float(lo)


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cpus, node, raw = local_cpus(dev)
    print("rank %d gpu %d numa_node %s local_cpulist %s affinity now %d cpus" % (rank, local, node, raw, len(os.sched_getaffinity(0))), flush=True)
    for mb in (22, 64):
        agg, lo = measure(dev, mb, 50, world)
        if rank == 0:
            print("default placement   %3d MB x %d ranks: aggregate %.1f GB/s, slowest rank %.1f GB/s" % (mb, world, agg, lo), flush=True)
    if cpus:
        os.sched_setaffinity(0, cpus)
        for mb in (22, 64):
            agg, lo = measure(dev, mb, 50, world)
            if rank == 0:
                print("bound to local cpus %3d MB x %d ranks: aggregate %.1f GB/s, slowest rank %.1f GB/s" % (mb, world, agg, lo), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
