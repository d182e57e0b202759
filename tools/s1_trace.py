#!/usr/bin/env python3
"""ROUND-1 TOOL (FP32 stage-1 kernel), kept for the record of the measurements cited in profiles/README.md: written against round 1's bench.py, it no longer runs; the tensor-core stage 1 has tools/s1t_trace.py.
Per-CTA phase timeline of the stage-1 kernel (debug build: SDRPP_EXTRA_NVCC=-DSDRPP_S1_TRACE, library passed
through SDRPP_CUDA_LIB). Prints, per launch of the last block, the mean duration of each phase and the fraction of
time both CTAs of an SM spend outside the main loop at once."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sdrpp_b200 import cuda  # noqa: E402

cuda.init(0)
fe = cuda.Frontend(bench.SR, fft_size=0, fft_rate=bench.SR / bench.FFT_N, fft_window=cuda.WIN_BH4, max_block=bench.BLOCK)
vf = bench.vfo_list()
only = os.environ.get("TRACE_CLASS")
if only is not None:
    vf = [v for i, v in enumerate(vf) if i % 2 == int(only)]
for v in vf:
    fe.add_vfo(*v)
blocks = bench.make_blocks(4)
fe.set_readback(False)
for i in range(6):
    fe.submit(cuda.FMT_CF32, blocks[i % 4])
    fe.wait()
L = cuda.lib()
rows = 8192
buf = np.zeros((rows, 6), dtype=np.int64)
L.sdrpp_cuda_debug_s1_trace.argtypes = [C.c_void_p, C.c_int]
rc = L.sdrpp_cuda_debug_s1_trace(buf.ctypes.data, rows)
assert rc == 0, rc
t = buf[buf[:, 0] > 0]
t0 = t[:, 0].min()
print("CTAs traced:", len(t), "span us:", (t[:, 4].max() - t0) / 1e3)
names = ["prologue (TMA wait)", "main loop", "rotate+parts+sync", "combine+store"]
for k in range(4):
    d = (t[:, k + 1] - t[:, k]) / 1e3
    print(f"  {names[k]:22s} mean {d.mean():7.2f} us  p10 {np.percentile(d, 10):7.2f}  p90 {np.percentile(d, 90):7.2f}")
tot = (t[:, 4] - t[:, 0]) / 1e3
print(f"  CTA total mean {tot.mean():.2f} us")
# per-SM overlap analysis
res = 0.05
for sm in sorted(set(t[:, 5]))[:3]:
    c = t[t[:, 5] == sm]
    c = c[np.argsort(c[:, 0])]
    print("SM", sm, "CTAs", len(c))
    for r in c[:8]:
        print("    start %8.2f  main %8.2f..%8.2f  end %8.2f" % tuple((r[[0, 1, 2, 4]] - t0) / 1e3))
busy_frac = []
for sm in set(t[:, 5]):
    c = t[t[:, 5] == sm]
    lo, hi = c[:, 0].min(), c[:, 4].max()
    grid = np.arange(lo, hi, 50)
    inmain = np.zeros(len(grid), dtype=np.int32)
    for r in c:
        inmain += ((grid >= r[1]) & (grid < r[2])).astype(np.int32)
    busy_frac.append([(inmain == 0).mean(), (inmain == 1).mean(), (inmain >= 2).mean()])
bf = np.array(busy_frac).mean(axis=0)
print("fraction of SM time with 0 / 1 / 2 CTAs in the main loop: %.3f / %.3f / %.3f" % tuple(bf))
