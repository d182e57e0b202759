#!/usr/bin/env python3
"""Host time of one submit (cfg5, pinned host blocks, readback on), with blocks replayed as CUDA graphs and command by
command, and the e2e rate of a SHORT run (20 steps after the graph warm-up) beside a long one: what the driver's
20-step bench sees."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdrpp_b200 import cuda, workloads  # noqa: E402

w = workloads.config(5)
cuda.init(0)
host = w.make_blocks(8)
blk_bytes = w.block * w.bytes_per_sample
pin = [cuda.PinnedArray((blk_bytes,), np.uint8) for _ in range(8)]
for j, p in enumerate(pin):
    p.array[:] = host[j].view(np.uint8)
for graphs in (1, 0):
    fe = cuda.Frontend(w.sr, decim_ratio=w.decim, fft_size=w.fft_size, fft_rate=w.fft_rate, fft_window=w.fft_window, max_block=w.block)
    fe.set_graphs(graphs)
    for v in w.vfos:
        fe.add_vfo(*v)
    for i in range(200):
        fe.submit(w.fmt, pin[i % 8], w.block); fe.wait()
    g0 = fe.graph_stats()
    for N in (20, 20, 400):
        ts = tw = 0.0
        t00 = time.perf_counter()
        for i in range(4):
            fe.submit(w.fmt, pin[i % 8], w.block)
        for i in range(4, N):
            t0 = time.perf_counter()
            fe.submit(w.fmt, pin[i % 8], w.block)
            t1 = time.perf_counter()
            fe.wait()
            t2 = time.perf_counter()
            ts += t1 - t0; tw += t2 - t1
        for i in range(4):
            fe.wait()
        tot = time.perf_counter() - t00
        g1 = fe.graph_stats()
        print(f"graphs={graphs} steps={N:4d}: {N * w.block / tot / 1e6:7.1f} MS/s, {tot / N * 1e6:6.1f} us per step; host time in submit {ts / (N - 4) * 1e6:5.1f} us, "
              f"in wait {tw / (N - 4) * 1e6:5.1f} us; graphs instantiated during the run: {g1['graphs_instantiated'] - g0['graphs_instantiated']}", flush=True)
        g0 = g1
    fe.close()
