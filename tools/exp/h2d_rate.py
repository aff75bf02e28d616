#!/usr/bin/env python
"""Host->device copy rate from regular pinned memory vs write-combined pinned memory (cudaHostAllocWriteCombined), one rank per
GPU, all ranks copying at the same time.  Is the e2e number at N = 8 bound by the host side of the copies?

    python tools/exp/h2d_rate.py                                         # one GPU
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 tools/exp/h2d_rate.py
"""
import ctypes
import os

import torch
import torch.distributed as dist
from cuda.bindings import runtime as cudart

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import pynvml
    pynvml.nvmlInit()
    aff = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local), (os.cpu_count() + 63) // 64)
    cpus = [i for i in range(os.cpu_count()) if aff[i // 64] >> (i % 64) & 1]
    if cpus:
        os.sched_setaffinity(0, cpus)
except Exception:                                             # noqa: BLE001
    pass

CH, NCH = 12 << 20, 32                                        # 12 MiB pieces (a 2^18-sample micro-batch is 12.6 MB), 384 MiB per pass
total = CH * NCH
dst = torch.empty(total, dtype=torch.uint8, device="cuda")


def host(kind):
    if kind == "pinned":
        t = torch.empty(total, dtype=torch.uint8).pin_memory()
        t.fill_(1)
        return t
    err, ptr = cudart.cudaHostAlloc(total, cudart.cudaHostAllocWriteCombined)
    assert int(err) == 0, err
    t = torch.frombuffer((ctypes.c_uint8 * total).from_address(int(ptr)), dtype=torch.uint8)
    t.fill_(1)
    return t


def rate(src, streams):
    ss = [torch.cuda.Stream() for _ in range(streams)]
    best = 0.0
    for rep in range(4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in ss:
            s.wait_stream(torch.cuda.current_stream())
        for i in range(NCH):
            with torch.cuda.stream(ss[i % streams]):
                dst[i * CH:(i + 1) * CH].copy_(src[i * CH:(i + 1) * CH], non_blocking=True)
        for s in ss:
            torch.cuda.current_stream().wait_stream(s)
        b.record()
        torch.cuda.synchronize()
        best = max(best, total / a.elapsed_time(b) / 1e6)
    t = torch.tensor([best], device="cuda")
    if world > 1:
        dist.all_reduce(t)
    return best, float(t)


for kind in ("pinned", "write_combined"):
    src = host(kind)
    for streams in (1, 2):
        mine, agg = rate(src, streams)
        if rank == 0:
            print("%-15s streams=%d  rank0 %.1f GB/s  aggregate over %d ranks %.1f GB/s" % (kind, streams, mine, world, agg), flush=True)
    del src
if world > 1:
    dist.destroy_process_group()
