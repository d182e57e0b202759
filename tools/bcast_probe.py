#!/usr/bin/env python3
"""How long does one NCCL broadcast of a raw block (4.9 MB) take on this box, back to back and alone? Launch with torchrun;
rank 0 prints. The multi-GPU step of bench.py cannot be shorter than this (one broadcast per block on a stream of its own)."""
import os
import sys

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for nbytes in (614400 * 8, 614400 * 2):
    buf = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(20):
        dist.broadcast(buf, 0)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for _ in range(n):
        dist.broadcast(buf, 0)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    t = torch.tensor([us], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"world {world}: broadcast of {nbytes / 1e6:.2f} MB back to back: {t.item():.1f} us each = {nbytes / t.item() / 1e3:.1f} GB/s  "
              f"[NCCL_PROTO={os.environ.get('NCCL_PROTO')} NCCL_ALGO={os.environ.get('NCCL_ALGO')} NCCL_MIN_NCHANNELS={os.environ.get('NCCL_MIN_NCHANNELS')}]", flush=True)
dist.destroy_process_group()
