#!/usr/bin/env python
"""NCCL all-to-all ceiling of the box (one process per GPU under torchrun): every rank sends `mb` MiB to every other
rank; reports GB/s sent per GPU (= per direction) for the best of a few repetitions.  Context for the `nvlink` object of
bench.py: the exchange of the fused substep moves the same pattern with the copy engines."""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for mb in (64, 512):
    n = mb * 2**20 // 8
    send = torch.ones(world * n, dtype=torch.float64, device="cuda")
    recv = torch.empty_like(send)
    best = 1e9
    for it in range(6):
        torch.cuda.synchronize()
        dist.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        dist.all_to_all_single(recv, send)
        t1.record()
        torch.cuda.synchronize()
        if it >= 2:
            best = min(best, t0.elapsed_time(t1))
    t = torch.tensor([best], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sent = (world - 1) * n * 8
        print(f"nccl all_to_all_single {world} ranks, {mb} MiB per peer: {t.item():.3f} ms, {sent / (t.item() * 1e-3) / 1e9:.1f} GB/s per direction per GPU", flush=True)
dist.destroy_process_group()
