"""ROUND-1 TOOL, kept for the record of the measurements cited in profiles/README.md: it was written against round 1's bench.py (torch.distributed broadcast in the Python loop) and no longer runs -- the broadcast lives inside the library now (csrc/comm.cu); use tools/bcast_probe.py and bench.py --gpus N.
Host-time breakdown of bench.py's multi-GPU step (broadcast of the IQ block + submit) and of cheaper variants.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/mgpu_probe.py [steps]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import bench                              # noqa: E402
from sdrpp_b200 import cuda, shard        # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    cuda.init(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    os.environ.setdefault("SDRPP_RESERVE_SMS", "8")
    vf_all = bench.vfo_list()
    costs = [shard.vfo_cost(bench.SR, v[0], v[1], cuda.design_resampler, cuda.design_decim_plan) for v in vf_all]
    mine = [vf_all[i] for i in shard.shard_vfos(costs, world, None)[rank]]
    fe = cuda.Frontend(bench.SR, fft_size=bench.FFT_N if rank == 0 else 0, fft_rate=bench.SR / bench.FFT_N, fft_window=cuda.WIN_BH4, max_block=bench.BLOCK)
    for v in mine:
        fe.add_vfo(*v)
    fe.set_readback(False)
    st = torch.cuda.ExternalStream(fe.stream, device=dev)
    NB = 8
    d_blocks = torch.empty((NB, bench.BLOCK, 2), dtype=torch.float32, device=dev)
    if rank == 0:
        d_blocks.copy_(torch.from_numpy(bench.make_blocks(NB).view(np.float32).reshape(NB, bench.BLOCK, 2)))
    nstage = 3
    d_stage = [torch.empty((bench.BLOCK, 2), dtype=torch.float32, device=dev) for _ in range(nstage)]
    bc = torch.cuda.Stream(device=dev)

    def run(variant):
        consumed = [None] * nstage
        ready = [torch.cuda.Event() for _ in range(nstage)]
        cons = [torch.cuda.Event() for _ in range(nstage)]
        used = [False] * nstage
        t = np.zeros(5)

        def step(i):
            blk = d_blocks[i % NB]
            if variant == 0:          # bench.py as it is
                k = i % 2
                buf = d_stage[k]
                cur = torch.cuda.current_stream()
                a = time.perf_counter()
                if consumed[k] is not None:
                    cur.wait_event(consumed[k])
                if rank == 0:
                    buf.copy_(blk, non_blocking=True)
                b = time.perf_counter()
                dist.broadcast(buf, src=0)
                c = time.perf_counter()
                ev = torch.cuda.Event()
                ev.record(cur)
                st.wait_event(ev)
                d = time.perf_counter()
                fe.submit_device(cuda.FMT_CF32, buf.data_ptr(), bench.BLOCK)
                e = time.perf_counter()
                consumed[k] = torch.cuda.Event()
                consumed[k].record(st)
                f = time.perf_counter()
            else:                      # broadcast on a side stream, straight out of the source block on rank 0, reused events
                k = i % nstage
                buf = d_stage[k]
                a = time.perf_counter()
                with torch.cuda.stream(bc):
                    if used[k]:
                        bc.wait_event(cons[k])
                    b = time.perf_counter()
                    if rank == 0 and variant == 1:
                        buf.copy_(blk, non_blocking=True)
                    src = blk if (rank == 0 and variant == 2) else buf
                    dist.broadcast(src, src=0)
                    c = time.perf_counter()
                    ready[k].record(bc)
                st.wait_event(ready[k])
                d = time.perf_counter()
                fe.submit_device(cuda.FMT_CF32, src.data_ptr(), bench.BLOCK)
                e = time.perf_counter()
                cons[k].record(st)
                used[k] = True
                f = time.perf_counter()
            t[:] += (b - a, c - b, d - c, e - d, f - e)

        for i in range(10):
            step(i)
        dist.barrier(); torch.cuda.synchronize()
        t[:] = 0
        t0 = time.perf_counter()
        for i in range(steps):
            step(10 + i)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dist.barrier()
        us = 1e6 * t / steps
        print(f"rank {rank} variant {variant}: {1e6 * dt / steps:6.1f} us per step = {bench.BLOCK / (dt / steps) / 1e6:7.0f} MS/s | host: pre {us[0]:5.1f} "
              f"bcast {us[1]:5.1f} events {us[2]:5.1f} submit {us[3]:5.1f} post {us[4]:5.1f}", flush=True)

    def run_threaded(from_source):
        """Broadcasts issued by a producer thread on a side stream, submits by the main thread: the two halves of the
        per-step host work (NCCL enqueue ~25-30 us, submit ~30-35 us) overlap; both release the GIL inside."""
        import queue
        import threading
        ready = [torch.cuda.Event() for _ in range(nstage)]
        cons = [torch.cuda.Event() for _ in range(nstage)]
        free = [threading.Semaphore(1) for _ in range(nstage)]   # host-side: cons[k] has been recorded for the previous use
        q = queue.Queue()

        def producer(first, count):
            torch.cuda.set_device(lr)
            with torch.cuda.stream(bc):
                for i in range(first, first + count):
                    k = i % nstage
                    free[k].acquire()
                    if i >= nstage:
                        bc.wait_event(cons[k])
                    blk = d_blocks[i % NB]
                    buf = d_stage[k]
                    if rank == 0 and not from_source:
                        buf.copy_(blk, non_blocking=True)
                    src = blk if (rank == 0 and from_source) else buf
                    dist.broadcast(src, src=0)
                    ready[k].record(bc)
                    q.put((k, src))

        def consume(count):
            for _ in range(count):
                k, src = q.get()
                st.wait_event(ready[k])
                fe.submit_device(cuda.FMT_CF32, src.data_ptr(), bench.BLOCK)
                cons[k].record(st)
                free[k].release()

        th = threading.Thread(target=producer, args=(0, 12))
        th.start(); consume(12); th.join()
        dist.barrier(); torch.cuda.synchronize()
        # counters restart at a multiple of nstage so that slot parity continues
        t0 = time.perf_counter()
        th = threading.Thread(target=producer, args=(12, steps))
        th.start(); consume(steps); th.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dist.barrier()
        print(f"rank {rank} threaded from_source={int(from_source)}: {1e6 * dt / steps:6.1f} us per step = {bench.BLOCK / (dt / steps) / 1e6:7.0f} MS/s", flush=True)

    for v in (0, 2):
        run(v)
    run_threaded(False)
    run_threaded(True)
    run_threaded(True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
