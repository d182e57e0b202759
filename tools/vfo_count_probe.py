"""Step time of the device-resident loop (bench.py's `value` loop) against the number of VFOs on one GPU: tells
whether the step is bound by the channelizer kernels or by something that does not scale with the VFO set.
  python tools/vfo_count_probe.py [steps] [nvfo ...]        (with one nvfo and steps <= 12 it is a good ncu target)"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from sdrpp_b200 import cuda, workloads    # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    counts = [int(a) for a in sys.argv[2:]] or [512, 256, 128, 64, 16]
    w = workloads.config(5)
    cuda.init(0)
    dev = torch.device("cuda:0")
    host = w.make_blocks(8)
    d_blocks = [torch.from_numpy(h.view(np.float32).reshape(-1, 2).copy()).to(dev) for h in host]
    for nv in counts:
        for fft in (True, False):
            fe = cuda.Frontend(w.sr, fft_size=w.fft_size if fft else 0, fft_rate=w.fft_rate, fft_window=w.fft_window, max_block=w.block)
            step = max(1, w.nvfo // nv)
            for v in w.vfos[::step][:nv]:
                fe.add_vfo(*v)
            fe.set_readback(False)
            for i in range(10):
                fe.submit_device(cuda.FMT_CF32, d_blocks[i % 8].data_ptr(), w.block)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            th = 0.0
            for i in range(steps):
                a = time.perf_counter()
                fe.submit_device(cuda.FMT_CF32, d_blocks[i % 8].data_ptr(), w.block)
                th += time.perf_counter() - a
            t_submit = time.perf_counter() - t0
            torch.cuda.synchronize()
            t1 = time.perf_counter() - t0
            fe.set_profiling(True)
            fam = []
            for i in range(6):
                fe.submit_device(cuda.FMT_CF32, d_blocks[i % 8].data_ptr(), w.block); fe.wait(); fam.append(fe.kernel_ms())
            fam = np.median(np.array(fam), axis=0) * 1e3
            print(f"vfos {nv:4d} spectrum {int(fft)}: {1e6 * t1 / steps:7.1f} us per step (host loop {1e6 * t_submit / steps:6.1f} us, inside submit "
                  f"{1e6 * th / steps:6.1f} us) launches/step {fe.launches / (steps + 16):.1f}  families us: ingest {fam[0]:.1f} spectrum {fam[1]:.1f} stage1 {fam[2]:.1f} tail {fam[3]:.1f}")
            fe.close()
            del fe


if __name__ == "__main__":
    main()
