import time, sys, os
import numpy as np
sys.path.insert(0, '/root/repo')
import bench
from sdrpp_b200 import cuda
cuda.init(0)
fe = cuda.Frontend(bench.SR, fft_size=bench.FFT_N, fft_rate=bench.SR / bench.FFT_N, fft_window=cuda.WIN_BH4, max_block=bench.BLOCK)
ids = [fe.add_vfo(*v) for v in bench.vfo_list()]
host = bench.make_blocks(4)
pin = [cuda.PinnedArray((bench.BLOCK,), np.complex64) for _ in range(4)]
for j, p in enumerate(pin):
    p.array[:] = host[j]
for rb in (True, False):
    fe.set_readback(rb)
    for i in range(10):
        fe.submit(cuda.FMT_CF32, pin[i % 4], bench.BLOCK); fe.wait()
    ts = tw = 0.0
    N = 300
    t00 = time.perf_counter()
    fe.submit(cuda.FMT_CF32, pin[0], bench.BLOCK)
    fe.submit(cuda.FMT_CF32, pin[1], bench.BLOCK)
    for i in range(2, N):
        t0 = time.perf_counter()
        fe.submit(cuda.FMT_CF32, pin[i % 4], bench.BLOCK)
        t1 = time.perf_counter()
        fe.wait()
        t2 = time.perf_counter()
        ts += t1 - t0; tw += t2 - t1
    fe.wait(); fe.wait()
    tot = time.perf_counter() - t00
    print(f"readback={rb}: per step {tot/N*1e6:.1f} us; host time in submit {ts/(N-2)*1e6:.1f} us, in wait {tw/(N-2)*1e6:.1f} us")
