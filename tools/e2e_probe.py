#!/usr/bin/env python3
"""Where does the end-to-end loop lose time against the device-resident loop? Times the same 3-deep submit/wait loop
(a) device-resident input, (b) device-resident input + an unrelated pinned H2D copy per step on a side stream,
(c) host input without readback, (d) host input with readback."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sdrpp_b200 import cuda  # noqa: E402

cuda.init(0)
torch.cuda.set_device(0)
fe = cuda.Frontend(bench.SR, fft_size=bench.FFT_N, fft_rate=bench.SR / bench.FFT_N, fft_window=cuda.WIN_BH4, max_block=bench.BLOCK)
ids = [fe.add_vfo(*v) for v in bench.vfo_list()]
host = bench.make_blocks(4)
pin = [cuda.PinnedArray((bench.BLOCK,), np.complex64) for _ in range(4)]
for j, p in enumerate(pin):
    p.array[:] = host[j]
dblk = torch.from_numpy(host.view(np.float32).reshape(4, bench.BLOCK, 2)).cuda()
side = torch.cuda.Stream()
hsrc = torch.empty((bench.BLOCK, 2), dtype=torch.float32).pin_memory()
hdst = torch.empty((bench.BLOCK, 2), dtype=torch.float32, device="cuda")
N = 400


def loop(mode):
    def sub(i):
        if mode in ("dev", "dev+h2d"):
            fe.submit_device(cuda.FMT_CF32, dblk[i % 4].data_ptr(), bench.BLOCK)
            if mode == "dev+h2d":
                with torch.cuda.stream(side):
                    hdst.copy_(hsrc, non_blocking=True)
        else:
            fe.submit(cuda.FMT_CF32, pin[i % 4], bench.BLOCK)
    for i in range(12):
        sub(i); fe.wait()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sub(0); sub(1)
    for i in range(2, N):
        sub(i); fe.wait()
    fe.wait(); fe.wait()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / N * 1e6


fe.set_readback(False)
print(f"device input, no readback      : {loop('dev'):7.1f} us per step")
print(f"device input + unrelated H2D   : {loop('dev+h2d'):7.1f} us per step")
print(f"host input, no readback        : {loop('host'):7.1f} us per step")
fe.set_readback(True)
print(f"device input, readback         : {loop('dev'):7.1f} us per step")
print(f"host input, readback           : {loop('host'):7.1f} us per step")

# which kernel family slows down while a host-to-device copy is running beside it?
fe.set_readback(False)
fe.set_profiling(True)
for bg in (False, True):
    fam = np.zeros(4)
    for i in range(40):
        if bg:
            with torch.cuda.stream(side):
                for _ in range(3):
                    hdst.copy_(hsrc, non_blocking=True)
        fe.submit_device(cuda.FMT_CF32, dblk[i % 4].data_ptr(), bench.BLOCK)
        fe.wait()
        if i >= 8:
            fam += np.array(fe.kernel_ms())
        torch.cuda.synchronize()
    print(("with H2D beside   " if bg else "alone             "), "ingest %.1f  spectrum %.1f  stage1 %.1f  tail %.1f us" % tuple(fam / 32 * 1e3))
