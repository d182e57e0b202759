"""Step time of the device-resident loop (bench.py's `value` loop) against the number of VFOs on one GPU: tells
whether the step is bound by the channelizer kernels or by something that does not scale with the VFO set.
  python tools/vfo_count_probe.py [steps]"""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench                              # noqa: E402
from sdrpp_b200 import cuda               # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 600
    cuda.init(0)
    dev = torch.device("cuda:0")
    host = bench.make_blocks(8)
    d_blocks = [torch.from_numpy(h.view("float32").reshape(-1, 2).copy()).to(dev) for h in host]
    for nv, fft in ((512, True), (256, True), (128, True), (512, False), (16, True)):
        fe = cuda.Frontend(bench.SR, fft_size=bench.FFT_N if fft else 0, fft_rate=bench.SR / bench.FFT_N, fft_window=cuda.WIN_BH4, max_block=bench.BLOCK)
        for v in bench.vfo_list()[:nv]:
            fe.add_vfo(*v)
        fe.set_readback(False)
        for i in range(10):
            fe.submit_device(cuda.FMT_CF32, d_blocks[i % 8].data_ptr(), bench.BLOCK)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        th = 0.0
        for i in range(steps):
            a = time.perf_counter()
            fe.submit_device(cuda.FMT_CF32, d_blocks[i % 8].data_ptr(), bench.BLOCK)
            th += time.perf_counter() - a
        t_submit = time.perf_counter() - t0
        torch.cuda.synchronize()
        t1 = time.perf_counter() - t0
        print(f"vfos {nv:4d} spectrum {int(fft)}: {1e6 * t1 / steps:7.1f} us per step "
              f"(host loop {1e6 * t_submit / steps:6.1f} us, inside submit {1e6 * th / steps:6.1f} us) launches/step {fe.launches / (steps + 10):.1f}")
        fe.close() if hasattr(fe, "close") else None
        del fe


if __name__ == "__main__":
    main()
