import os, sys, time
import numpy as np
sys.path.insert(0, "/root/repo")
from sdrpp_b200 import cuda, workloads
w = workloads.config(5)
cuda.init(0)
host = w.make_blocks(8)
blk_bytes = w.block * w.bytes_per_sample
pin = [cuda.PinnedArray((blk_bytes,), np.uint8) for _ in range(8)]
for j, p in enumerate(pin):
    p.array[:] = host[j].view(np.uint8)
fe = cuda.Frontend(w.sr, decim_ratio=w.decim, fft_size=w.fft_size, fft_rate=w.fft_rate, fft_window=w.fft_window, max_block=w.block)
for v in w.vfos:
    fe.add_vfo(*v)
for i in range(200):
    fe.submit(w.fmt, pin[i % 8], w.block); fe.wait()
for rep in range(2):
    N = 20
    ev = []
    t00 = time.perf_counter()
    for i in range(4):
        fe.submit(w.fmt, pin[i % 8], w.block); ev.append(("s%d" % i, time.perf_counter() - t00))
    for i in range(4, N):
        fe.submit(w.fmt, pin[i % 8], w.block); ev.append(("s%d" % i, time.perf_counter() - t00))
        fe.wait(); ev.append(("w%d" % (i - 4), time.perf_counter() - t00))
    for i in range(4):
        fe.wait(); ev.append(("w%d" % (N - 4 + i), time.perf_counter() - t00))
    print(" ".join(f"{n}:{t*1e6:.0f}" for n, t in ev))
fe.close()
