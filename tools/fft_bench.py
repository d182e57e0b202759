#!/usr/bin/env python3
"""Batched spectrum throughput of sdrpp_cuda_spectrum_device (used for ncu captures of the FFT kernels).
   python tools/fft_bench.py [log2N] [frames] [reps] [bufs]     bufs > 1: the calls cycle through that many inputs (more than L2)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdrpp_b200 import cuda  # noqa: E402

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
F = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
bufs = int(sys.argv[4]) if len(sys.argv) > 4 else 1
N = 1 << lg
cuda.init(0)
x = torch.randn((F * N, 2), dtype=torch.float32, device="cuda") * 0.1
xs = [x] + [x.clone() for _ in range(bufs - 1)]
rows = torch.empty((F, N), dtype=torch.float32, device="cuda")
win = cuda.design_window(cuda.WIN_BH4, N)
st = torch.cuda.Stream()
for _ in range(3):
    cuda.spectrum_device(N, N, F, N, x.data_ptr(), win, rows.data_ptr(), st.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for r in range(reps):
    cuda.spectrum_device(N, N, F, N, xs[r % bufs].data_ptr(), None, rows.data_ptr(), st.cuda_stream)  # cached window
e1.record(st)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"N=2^{lg} frames={F} bufs={bufs}: {ms:.4f} ms/call, {F * N / ms / 1e6:.1f} GS/s, {12.0 * F * N / ms / 1e6:.0f} GB/s algorithmic")
