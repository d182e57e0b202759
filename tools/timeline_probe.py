#!/usr/bin/env python3
"""Device timeline of the real multi-stream pipeline (cfg5, device-resident loop): CUDA timing events recorded on the streams
the kernels run on (SDRPP_TIMELINE=1; blocks run command by command, no graphs). Prints, for the last blocks, when the main
stream's ingest/split and stage 1, the spectrum and the tail's wide stage and tail kernel started and ended."""
import ctypes as C
import os
import sys

import numpy as np
import torch

os.environ["SDRPP_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdrpp_b200 import cuda, workloads  # noqa: E402

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
dblk = torch.from_numpy(host.view(np.uint8).reshape(NB, blk_bytes)).cuda()
fe.set_readback(False)
for i in range(120):
    fe.submit_device(w.fmt, dblk[i % NB].data_ptr(), w.block)
torch.cuda.synchronize()
out = (C.c_float * (16 * 9))()
L = cuda.lib()
L.sdrpp_cuda_debug_timeline.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
n = L.sdrpp_cuda_debug_timeline(fe.h, out, 16)
a = np.array(out[:n * 9]).reshape(n, 9)
print("times in us relative to the oldest block's start; per block: main[start split_done s1_done] fft[start done] tail[start wide_done tail_done]")
for r in a:
    t = r[1:] * 1e3
    print(f"blk {int(r[0]):4d}: main {t[0]:7.1f} {t[1]:7.1f} {t[2]:7.1f} | fft {t[3]:7.1f} {t[4]:7.1f} | tail {t[5]:7.1f} {t[6]:7.1f} {t[7]:7.1f}"
          f"   [split {t[1]-t[0]:5.1f} s1 {t[2]-t[1]:5.1f} fft {t[4]-t[3]:5.1f} wide {t[6]-t[5]:5.1f} tail {t[7]-t[6]:5.1f}]")
d = np.diff(a[:, 3]) * 1e3
print("step (s1 done to s1 done): median %.1f us" % float(np.median(d)))
fe.close()
