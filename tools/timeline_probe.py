#!/usr/bin/env python3
"""Device timeline of the real multi-stream pipeline (cfg5): CUDA timing events recorded on the streams the kernels run on
(SDRPP_TIMELINE=1; blocks run command by command, no graphs). Prints, per block, when the host-to-device copy, the main
stream's ingest/split and stage 1, the spectrum, the tail's wide stage and tail kernel and the result copies ended.
  python tools/timeline_probe.py            device-resident blocks, steady state (last 16 of 120 blocks)
  python tools/timeline_probe.py short      pinned host blocks, readback, 20 blocks from an EMPTY pipeline, four ahead
                                            (what the driver's 20-step end-to-end leg looks like)"""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["SDRPP_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdrpp_b200 import cuda, workloads  # noqa: E402

short = len(sys.argv) > 1 and sys.argv[1] == "short"
w = workloads.config(int(os.environ.get("PROBE_CONFIG", "5")))
cuda.init(0)
fe = cuda.Frontend(w.sr, decim_ratio=w.decim, fft_size=w.fft_size, fft_rate=w.fft_rate, fft_window=w.fft_window, max_block=w.block)
nv = int(os.environ.get("PROBE_VFOS", str(w.nvfo)))
step = max(1, w.nvfo // nv)
for v in w.vfos[::step][:nv]:
    fe.add_vfo(*v)
NB = 8
host = w.make_blocks(NB)
blk_bytes = w.block * w.bytes_per_sample
if short:
    pin = [cuda.PinnedArray((blk_bytes,), np.uint8) for _ in range(NB)]
    for j, p in enumerate(pin):
        p.array[:] = host[j].view(np.uint8)
    for i in range(40):
        fe.submit(w.fmt, pin[i % NB], w.block); fe.wait()
    torch.cuda.synchronize()
    N = 20
    for i in range(4):
        fe.submit(w.fmt, pin[i % NB], w.block)
    for i in range(4, N):
        fe.submit(w.fmt, pin[i % NB], w.block); fe.wait()
    for i in range(4):
        fe.wait()
else:
    dblk = torch.from_numpy(host.view(np.uint8).reshape(NB, blk_bytes)).cuda()
    fe.set_readback(False)
    for i in range(120):
        fe.submit_device(w.fmt, dblk[i % NB].data_ptr(), w.block)
torch.cuda.synchronize()
out = (C.c_float * (32 * 11))()
L = cuda.lib()
L.sdrpp_cuda_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
n = L.sdrpp_cuda_debug_timeline(fe.h, out, 32)
a = np.array(out[:n * 11]).reshape(n, 11)
a = a[-20:] if short else a[-16:]
t00 = a[0, 1]
print("us relative to the first block shown; main[start split_done s1_done] fft[start done] tail[start wide_done tail_done] h2d_done results_on_host")
for r in a:
    t = (r[1:] - t00) * 1e3
    print(f"blk {int(r[0]):4d}: h2d {t[8]:7.1f} | main {t[0]:7.1f} {t[1]:7.1f} {t[2]:7.1f} | fft {t[3]:7.1f} {t[4]:7.1f} | tail {t[5]:7.1f} {t[6]:7.1f} {t[7]:7.1f} | out {t[9]:7.1f}"
          f"   [split {t[1]-t[0]:5.1f} s1 {t[2]-t[1]:5.1f} fft {t[4]-t[3]:5.1f} wide {t[6]-t[5]:5.1f} tail {t[7]-t[6]:5.1f}]")
d = np.diff(a[:, 3]) * 1e3
print("step (s1 done to s1 done): median %.1f us" % float(np.median(d)))
fe.close()
