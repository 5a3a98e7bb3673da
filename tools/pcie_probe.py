#!/usr/bin/env python3
"""Concurrent pinned D2H bandwidth per rank, with and without binding the process to the CPUs local to its GPU.
torchrun --nproc-per-node N tools/pcie_probe.py"""
import os, time
import torch, torch.distributed as dist
import pynvml

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(local)
ncpu = os.cpu_count()
words = (ncpu + 63) // 64
aff = pynvml.nvmlDeviceGetCpuAffinity(h, words)
cpus = [i for i in range(ncpu) if (aff[i // 64] >> (i % 64)) & 1]
bdf = pynvml.nvmlDeviceGetPciInfo(h).busId
try:
    numa = open(f"/sys/bus/pci/devices/{bdf.lower()[4:] if bdf.startswith('0000') else bdf.lower()}/numa_node").read().strip()
except Exception as e:
    numa = f"? ({e})"

def measure(tag):
    n = 200 << 20
    dev = torch.empty(n, dtype=torch.uint8, device="cuda")
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.copy_(dev); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        host.copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = 10 * n / dt / 1e9
    out = torch.tensor([gbs], device="cuda")
    if world > 1:
        lst = [torch.zeros(1, device="cuda") for _ in range(world)]
        dist.all_gather(lst, out)
        vals = [float(x) for x in lst]
    else:
        vals = [gbs]
    if rank == 0:
        print(tag, " ".join(f"{v:.1f}" for v in vals), "GB/s per rank; sum", f"{sum(vals):.1f}", flush=True)

print(f"rank {rank}: gpu {local} bdf {bdf} numa {numa} local cpus {cpus[:4]}..{cpus[-1] if cpus else None} ({len(cpus)}) current affinity {len(os.sched_getaffinity(0))}", flush=True)
measure("unbound:")
if cpus:
    os.sched_setaffinity(0, cpus)
measure("bound to the GPU's CPUs:")
if world > 1:
    dist.destroy_process_group()
