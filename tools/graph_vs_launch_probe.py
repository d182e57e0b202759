#!/usr/bin/env python3
"""Do kernel launches wait for their commands while a host-to-device copy occupies the link, and does a CUDA graph
avoid it? 12 dependent ~8 us kernels per step, eager and as one graph launch, alone and with a 4.9 MB pinned H2D copy
per step on a side stream."""
import time

import torch

torch.cuda.set_device(0)
x = torch.randn(8 << 20, device="cuda")
ys = [torch.empty_like(x) for _ in range(2)]
side = torch.cuda.Stream()
hsrc = torch.empty(4915200 // 4, dtype=torch.float32).pin_memory()
hdst = torch.empty(4915200 // 4, dtype=torch.float32, device="cuda")


def step():
    a = x
    for i in range(12):
        torch.mul(a, 1.0001, out=ys[i & 1])
        a = ys[i & 1]


for _ in range(3):
    step()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()


def run(fn, copies, n=300):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        if copies:
            with torch.cuda.stream(side):
                hdst.copy_(hsrc, non_blocking=True)
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for name, fn in (("eager, 12 launches", step), ("one graph launch  ", g.replay)):
    print(f"{name}: alone {run(fn, False):7.1f} us per step, with H2D beside {run(fn, True):7.1f} us per step")
