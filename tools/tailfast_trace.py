"""Per-phase cycle trace of tail_fast_kernel (CTA 0, thread 0). Build the library with
SDRPP_EXTRA_NVCC=-DSDRPP_TAILFAST_TRACE into a separate .so and point SDRPP_CUDA_LIB at it:
  SDRPP_EXTRA_NVCC=-DSDRPP_TAILFAST_TRACE python -m sdrpp_b200.build --force && python tools/tailfast_trace.py [nvfo]"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from sdrpp_b200 import cuda, workloads    # noqa: E402

nv = int(sys.argv[1]) if len(sys.argv) > 1 else 16
w = workloads.config(5)
cuda.init(0)
host = w.make_blocks(4)
dev = torch.device("cuda:0")
d = [torch.from_numpy(h.view(np.float32).reshape(-1, 2).copy()).to(dev) for h in host]
fe = cuda.Frontend(w.sr, max_block=w.block)
for v in w.vfos[::max(1, w.nvfo // nv)][:nv]:
    fe.add_vfo(*v)
fe.set_readback(False)
for i in range(6):
    fe.submit_device(cuda.FMT_CF32, d[i % 4].data_ptr(), w.block)
torch.cuda.synchronize()
out = (C.c_longlong * 16)()
cuda.lib().sdrpp_cuda_debug_tailfast_trace(out)
t = list(out)
names = ["entry", "layout", "loads issued", "loads landed + sync", "stage0", "stage1", "stage2", "stage3", "stage4", "stage5", "", "", "end"]
print("last FIR stage: setup->loop start", t[13] - t[7], "loop", t[14] - t[13], "reduce+store", t[15] - t[14], "final sync", t[8] - t[15])
prev = t[0]
for i in range(1, 13):
    if t[i] > 0:
        print(f"{names[i]:22s} +{t[i] - prev:7d} cycles   (total {t[i] - t[0]})")
        prev = t[i]
