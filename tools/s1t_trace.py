#!/usr/bin/env python3
"""Per-CTA cycle counters of the tensor-core stage-1 kernel (debug build: SDRPP_EXTRA_NVCC=-DSDRPP_S1T_TRACE).
Prints, for the last block, how long each role waited on each barrier."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdrpp_b200 import cuda, workloads  # noqa: E402

cuda.init(0)
w = workloads.config(int(os.environ.get("S1T_TRACE_CONFIG", "5")))
fe = cuda.Frontend(w.sr, fft_size=0, fft_rate=w.fft_rate, fft_window=w.fft_window, max_block=w.block)
for v in w.vfos:
    fe.add_vfo(*v)
blocks = w.make_blocks(4)
fe.set_readback(False)
for i in range(6):
    fe.submit(cuda.FMT_CF32, blocks[i % 4])
    fe.wait()
L = cuda.lib()
buf = np.zeros((256, 24), dtype=np.int64)
L.sdrpp_cuda_debug_s1t_trace.argtypes = [C.c_void_p, C.c_int]
assert L.sdrpp_cuda_debug_s1t_trace(buf.ctypes.data, 256) == 0
t = buf[buf[:, 12] > 0]
print(f"kernel span (globaltimer, first CTA start -> last CTA end): {(t[:, 0].max() - t[:, 11].min()) / 1e3:.1f} us; "
      f"CTA starts spread over {(t[:, 11].max() - t[:, 11].min()) / 1e3:.1f} us; CTA ends spread over {(t[:, 0].max() - t[:, 0].min()) / 1e3:.1f} us")
fe.set_profiling(True)
for i in range(4):
    fe.submit(cuda.FMT_CF32, blocks[i % 4]); fe.wait()
print("family brackets (ms): ingest, spectrum, stage1, tail =", fe.kernel_ms())
names = {1: "kernel total", 2: "mma: wait tmem empty", 3: "mma: wait smem full", 4: "mma: issue hi chunks (16 MMA)", 8: "mma: issue lo chunks (8 MMA)", 5: "epi: wait tmem full", 6: "epi: load+sum phase",
         7: "epi: bar.sync", 9: "epi: tile total (after wait)", 16: "epi: scale load + shuffles", 17: "epi: store phase", 18: "epi: exact phase (every 8 tiles)", 10: "producer: wait smem empty", 14: "start -> B image landed", 15: "start -> first accumulator ready"}
for A in sorted(set(t[:, 13])):
    c = t[t[:, 13] == A]
    print(f"A={A}: {len(c)} CTAs, tiles per CTA {c[:, 12].min()}..{c[:, 12].max()}")
    for k, n in names.items():
        per = c[:, k] / c[:, 12]
        print(f"   {n:32s} total {c[:, k].mean():10.0f} cyc   per tile {per.mean():8.0f}")
