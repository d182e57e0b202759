#!/usr/bin/env python3
"""Where does the end-to-end loop lose time against the device-resident loop (cfg5)? The same submit / wait loop with four
blocks submitted ahead, in variants that add ONE ingredient of the end-to-end loop at a time:

  dev            device-resident blocks, no readback                      (bench.py's `value`)
  dev+h2d        + an unrelated 4.9 MB pinned host-to-device copy per step on a side stream
  dev+d2h        + an unrelated 3.6 MB device-to-host copy per step on a side stream
  dev+h2d+d2h    both
  dev,readback   device-resident blocks, results copied back
  host           pinned host blocks, no readback
  host,readback  the end-to-end loop                                       (bench.py's `e2e`)

and the per-family kernel times (profiling mode: one stream, CUDA events) alone and with the copies running beside them.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdrpp_b200 import cuda, workloads  # noqa: E402

w = workloads.config(int(os.environ.get("PROBE_CONFIG", "5")))
cuda.init(0)
torch.cuda.set_device(0)
fe = cuda.Frontend(w.sr, decim_ratio=w.decim, fft_size=w.fft_size, fft_rate=w.fft_rate, fft_window=w.fft_window, max_block=w.block)
ids = [fe.add_vfo(*v) for v in w.vfos]
NB = 8
host = w.make_blocks(NB)
blk_bytes = w.block * w.bytes_per_sample
pin = [cuda.PinnedArray((blk_bytes,), np.uint8) for _ in range(NB)]
for j, p in enumerate(pin):
    p.array[:] = host[j].view(np.uint8)
dblk = torch.from_numpy(host.view(np.uint8).reshape(NB, blk_bytes)).cuda()
side_h2d, side_d2h = torch.cuda.Stream(), torch.cuda.Stream()
hsrc = torch.empty((blk_bytes,), dtype=torch.uint8).pin_memory()
hdst = torch.empty((blk_bytes,), dtype=torch.uint8, device="cuda")
d2h_bytes = 3563520
dsrc2 = torch.empty((d2h_bytes,), dtype=torch.uint8, device="cuda")
hdst2 = torch.empty((d2h_bytes,), dtype=torch.uint8).pin_memory()
N = int(os.environ.get("PROBE_STEPS", "600"))
AHEAD = 4


def loop(mode):
    host_in = mode.startswith("host")

    def sub(i):
        if host_in:
            fe.submit(w.fmt, pin[i % NB], w.block)
        else:
            fe.submit_device(w.fmt, dblk[i % NB].data_ptr(), w.block)
        if "+h2d" in mode:
            with torch.cuda.stream(side_h2d):
                hdst.copy_(hsrc, non_blocking=True)
        if "+d2h" in mode:
            with torch.cuda.stream(side_d2h):
                hdst2.copy_(dsrc2, non_blocking=True)
    for i in range(200):
        sub(i); fe.wait()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(AHEAD):
        sub(i)
    for i in range(AHEAD, N):
        sub(i); fe.wait()
    for i in range(AHEAD):
        fe.wait()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / N * 1e6


for rb, modes in ((False, ["dev", "dev+h2d", "dev+d2h", "dev+h2d+d2h", "host"]), (True, ["dev", "host"])):
    fe.set_readback(rb)
    for m in modes:
        us = loop(m)
        print(f"{m + (',readback' if rb else ''):18s}: {us:7.1f} us per step  {w.block / us:8.1f} MS/s", flush=True)

# which kernel family slows down while copies run beside it?
fe.set_readback(False)
fe.set_profiling(True)
for bg in ("", "+h2d", "+d2h"):
    fam = np.zeros(4)
    cnt = 0
    for i in range(60):
        if "+h2d" in bg:
            with torch.cuda.stream(side_h2d):
                for _ in range(3):
                    hdst.copy_(hsrc, non_blocking=True)
        if "+d2h" in bg:
            with torch.cuda.stream(side_d2h):
                for _ in range(3):
                    hdst2.copy_(dsrc2, non_blocking=True)
        fe.submit_device(w.fmt, dblk[i % NB].data_ptr(), w.block)
        fe.wait()
        if i >= 12:
            fam += np.array(fe.kernel_ms()); cnt += 1
        torch.cuda.synchronize()
    print(f"profiling mode {bg or 'alone':6s}: ingest %.1f  spectrum %.1f  stage1 %.1f  tail %.1f us" % tuple(fam / cnt * 1e3), flush=True)

# The host link on its own: the 4.9 MB block copy back to back, alone and with the result-sized device-to-host copies of the
# end-to-end loop running the other way at the same time (PCIe is full duplex, but read requests and write data share the
# upstream direction: the bidirectional figure is the ceiling of an end-to-end number that reads every result back).
def link(bidir, n=200):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(side_h2d):
        e0.record()
        for _ in range(n):
            hdst.copy_(hsrc, non_blocking=True)
        e1.record()
    if bidir:
        with torch.cuda.stream(side_d2h):
            for _ in range(int(n * 1.5)):
                hdst2.copy_(dsrc2, non_blocking=True)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    return blk_bytes / ms / 1e6, ms * 1e3


for bidir in (False, True):
    gbs, us = link(bidir)
    print(f"link: 4.9 MB H2D {'with 3.6 MB D2H copies the other way' if bidir else 'alone':36s}: {gbs:6.1f} GB/s, {us:6.1f} us per block -> {w.block / us:7.1f} MS/s ceiling", flush=True)
fe.close()
