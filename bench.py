#!/usr/bin/env python3
"""bench.py -- sustained input MS/s of the SDR++ signal-path hot loop on B200.

Workload (BASELINE.json configs[4] + the 1M-point spectrum the metric is quoted on): 122.88 MS/s
complex64 IQ in blocks of 614,400 samples (sr/200), a saturated 1,048,576-point Blackman-Harris-4
spectrum (every sample enters one frame) and 512 VFOs alternating NFM (12.5 kHz -> 48 kS/s,
quadrature demod) and AM (12 kHz -> 24 kS/s, magnitude). A "step" is one IQ block through
conversion/ingest -> spectrum frames -> 512-VFO channelizer -> demod front ends.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1 (torchrun, one rank per GPU): the VFO set is sharded across ranks, every step the block is
broadcast from rank 0 over NCCL (NVLink) and the spectrum stays on rank 0.
"""
import argparse
import collections
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 122.88e6
BLOCK = 614400
FFT_N = 1 << 20
NVFO = 512
NFM = (48e3, 12.5e3, 1)   # outSR, bw, demod (quadrature)
AM = (24e3, 12e3, 2)      # outSR, bw, demod (magnitude)
WORKLOAD = ("122.88 MS/s cf32 IQ, blocks of 614400; saturated 1048576-pt Blackman-Harris-4 spectrum; "
            "512 VFOs alternating NFM 12.5k->48k (quadrature) / AM 12k->24k (magnitude)")
METRIC = "sustained input MS/s (1M-pt FFT + N-VFO channelizer); % of B200 HBM roofline"


def vfo_list():
    from sdrpp_b200 import synth
    offs = synth.vfo_grid(NVFO, SR)
    return [((NFM if i % 2 == 0 else AM)[0], (NFM if i % 2 == 0 else AM)[1], float(offs[i]), (NFM if i % 2 == 0 else AM)[2])
            for i in range(NVFO)]


def make_blocks(nblocks, seed=5):
    """Synthetic IQ: a few tones/carriers + white noise at -40 dBFS, complex64, nblocks x BLOCK."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = nblocks * BLOCK
    out = np.empty(n, dtype=np.complex64)
    offs = [v[2] for v in vfo_list()[::64]]
    chunk = 1 << 20
    sigma = np.float32(10.0 ** (-40.0 / 20.0) / np.sqrt(2.0))
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        t = (np.arange(s, s + m, dtype=np.float64)) / SR
        x = np.zeros(m, dtype=np.complex64)
        for i, f in enumerate(offs):
            ph = (2.0 * np.pi) * ((f * t) % 1.0)
            x += np.float32(0.05) * (np.cos(ph) + 1j * np.sin(ph)).astype(np.complex64)
        x += sigma * (rng.standard_normal(m, dtype=np.float32) + 1j * rng.standard_normal(m, dtype=np.float32))
        out[s:s + m] = x
    return out.reshape(nblocks, BLOCK)


# ---------------------------------------------------------------------------------------------
# algorithmic work per input sample
# ---------------------------------------------------------------------------------------------
def algorithmic_model(vfos):
    from sdrpp_b200 import cuda
    flops = 0.0
    out_bytes = 0.0
    s1_flops = 0.0
    for (osr, bw, _off, demod) in vfos:
        info, _ = cuda.design_resampler(SR, osr)
        f = 8.0  # NCO complex multiply + phase advance per input sample
        rate = 1.0
        stages = cuda.design_decim_plan(info["predec"]) if info["mode"] in (0, 1) else []
        for i, (d, taps) in enumerate(stages):
            c = 4.0 * len(taps) * rate / d   # 2 mul + 2 add per tap per output
            f += c
            if i == 0:
                s1_flops += 8.0 + c
            rate /= d
        if info["mode"] in (0, 2):
            f += 4.0 * info["tpp"] * (osr / SR)
        if bw != osr:
            f += 4.0 * int(3.8 * osr / (bw / 20.0)) * (osr / SR)
        flops += f
        out_bytes += (osr / SR) * (8 + (4 if demod else 0))
    return dict(flops_per_sample=flops, stage1_flops_per_sample=s1_flops, chan_bytes_per_sample=8.0 + out_bytes,
                fft_bytes_per_sample=12.0, fft_flops_per_sample=5.0 * 20 + 12)


# ---------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled every ~2 ms from a thread
    (the timed region is tens of milliseconds, too short for `nvidia-smi -lms`), nvidia-smi as a fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self.stop_flag = False
        self.th = None
        self.nvml = None

    def _poll(self):
        n = self.nvml
        h = n.nvmlDeviceGetHandleByIndex(self.idx)
        bits = {"hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self.stop_flag:
            try:
                self.samples.append(float(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for name, b in bits.items():
                    if r & b:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
        except Exception:
            self.nvml = None

    def stop(self):
        if self.nvml is not None and self.th is not None:
            self.stop_flag = True
            self.th.join(timeout=1.0)
            if self.samples:
                return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                        "samples": len(self.samples), "source": "nvml, 2 ms poll during the timed region"}
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip().splitlines()[0]
            p = [x.strip() for x in out.split(",")]
            reasons = [nm for nm, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]) if v.lower().startswith("active")]
            return {"sm_mhz": float(p[1]), "sm_max_mhz": float(p[2]), "reasons": reasons, "samples": 1, "source": "nvidia-smi right after the timed region"}
        except Exception:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}


# ---------------------------------------------------------------------------------------------
# CPU reference arm / baseline: the reference's own dsp/ headers (oracle/_ref, release flags)
# ---------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def cpu_reference(nblocks=2, sample_vfos=None, fft_frames=1):
    """Times the reference's CPU implementation of the path on a bounded sample of the workload:
    `sample_vfos` of the 512 VFOs (thread-per-VFO multiplexed on all host cores, as
    ref_bench_channelizer does) on nblocks blocks plus fft_frames 1M-point spectrum lines, scaled to
    the full VFO count. Returns (MS/s, info dict)."""
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    kind = "reference"
    if po.have_ref("fast"):
        lib = po.Ref("fast")
    elif po.have_ref(""):
        lib = po.Ref("")
    else:
        lib = None
    vf = vfo_list()
    if sample_vfos is None:
        sample_vfos = min(NVFO, 2 * cores)
    pick = [vf[(i * NVFO) // sample_vfos] for i in range(sample_vfos)]
    if "blocks" not in _CPU_CACHE:
        _CPU_CACHE["blocks"] = make_blocks(2, seed=5).reshape(-1)  # 1,228,800 samples >= one spectrum frame
    blocks = _CPU_CACHE["blocks"]
    if lib is not None:
        win = lib.window(po.WIN_BH4, FFT_N)
        # the block handed to every VFO thread is the first BLOCK samples; the spectrum thread reads FFT_N
        buf = np.ascontiguousarray(blocks[:max(BLOCK, FFT_N)])
        # ref_bench_channelizer(count=len(block)): pass exactly BLOCK for the VFOs by timing the two legs apart
        t_ch = lib.bench_channelizer(SR, pick, buf[:BLOCK], nblocks, cores)
        t_fft = lib.bench_channelizer(SR, pick[:1], buf, 0, 1, fft=(FFT_N, win, fft_frames))
    else:
        kind = "port"
        port = po.Port()
        t0 = time.perf_counter()
        for v in pick:
            o = port.rxvfo(SR, v[0], v[1], v[2])
            for _ in range(nblocks):
                o.process(blocks[:BLOCK])
        t_ch = time.perf_counter() - t0
        cores_used = 1
        t0 = time.perf_counter()
        win = port.window(po.WIN_BH4, FFT_N)
        for _ in range(fft_frames):
            port.spectrum(FFT_N, blocks[:FFT_N], win, want64=False)
        t_fft = time.perf_counter() - t0
        cores = cores_used
    # the channelizer workers and the spectrum thread run concurrently in the reference; the stream
    # advances at the pace of the slower consumer
    sec_per_sample_ch = (t_ch * (NVFO / float(sample_vfos))) / (nblocks * BLOCK)
    sec_per_sample_fft = t_fft / (fft_frames * FFT_N)
    msps = 1e-6 / max(sec_per_sample_ch, sec_per_sample_fft)
    info = dict(kind=kind, cores=cores,
                sample=f"{sample_vfos} of {NVFO} VFOs x {nblocks} blocks of {BLOCK} (scaled x{NVFO / sample_vfos:g}) on {cores} threads "
                       f"+ {fft_frames} x 1M-pt spectrum line on 1 thread; generic-VOLK shim, -O3 -ffast-math",
                channelizer_msps=1e-6 / sec_per_sample_ch, spectrum_msps=1e-6 / sec_per_sample_fft,
                seconds=t_ch + t_fft)
    return msps, info


def run_reference(args, rank, world):
    if rank != 0:
        return
    vals, info = [], None
    for i in range(args.warmup + args.steps):
        v, info = cpu_reference(nblocks=1)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * BLOCK / (value * 1e6), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "block": BLOCK, "vfos": NVFO, "fft": FFT_N},
        "cpu_baseline": {"value": value, "unit": "MS/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"]},
        "e2e": {"value": value, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from sdrpp_b200 import cuda
    from oracle import pyoracle as po  # only for the cpu_baseline leg below (rank 0, N = 1)

    if cuda.device_count() <= 0:
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    cuda.init(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.pop("NCCL_DEBUG", None)  # any level >= VERSION makes NCCL print a banner on stdout; rank 0 prints ONE JSON line
        dist.init_process_group("nccl", device_id=dev)

    if world > 1:
        # the stage-1 kernel is persistent (one CTA per SM): leave a few SMs to the NCCL broadcast of the next block
        os.environ.setdefault("SDRPP_RESERVE_SMS", "8")
    vf_all = vfo_list()
    # shard the VFO set across ranks, balanced by per-VFO cost (sdrpp_b200/shard.py); no data-path collective
    from sdrpp_b200 import shard
    costs = [shard.vfo_cost(SR, v[0], v[1], cuda.design_resampler, cuda.design_decim_plan) for v in vf_all]
    # rank 0 also ingests, broadcasts and runs the 1M-point spectrum (~25 us per step); at the measured
    # ~32 ns per cost unit of stage 1 (0.256 ms for 512 VFOs x 15.4) that is ~770 units of VFO work it cannot take
    base = [770.0] + [0.0] * (world - 1) if world > 1 else None
    mine = [vf_all[i] for i in shard.shard_vfos(costs, world, base)[rank]]
    with_fft = (rank == 0)
    fe = cuda.Frontend(SR, fft_size=FFT_N if with_fft else 0, fft_rate=SR / FFT_N, fft_window=cuda.WIN_BH4, max_block=BLOCK)
    ids = [fe.add_vfo(*v) for v in mine]
    st = torch.cuda.ExternalStream(fe.stream, device=dev)

    NB = args.input_blocks
    host = make_blocks(NB) if rank == 0 else None
    d_blocks = torch.empty((NB, BLOCK, 2), dtype=torch.float32, device=dev)  # > L2 (126 MB) when NB >= 32
    if rank == 0:
        d_blocks.copy_(torch.from_numpy(host.view(np.float32).reshape(NB, BLOCK, 2)))
    # Multi-GPU: the IQ stream is broadcast in buckets of KB consecutive blocks (one NCCL call per bucket, straight out
    # of the source buffer on the ingest rank, into one of two staging buckets elsewhere). The enqueue cost of a
    # collective (~25-30 us of host time through torch.distributed) is what bounded the loop with one broadcast per
    # block (tools/mgpu_probe.py); the bucket size is chosen for that latency, not for the link. Every block is
    # still submitted on its own.
    KB = shard.fit_bucket(args.bcast_blocks, NB) if world > 1 else 1
    if world > 1:
        d_stage = [torch.empty((KB, BLOCK, 2), dtype=torch.float32, device=dev) for _ in range(2)]
    torch.cuda.synchronize()

    consumed = [None, None]  # per staging bucket: event after which the front end no longer reads it
    pos = [0]                # blocks submitted in the current phase (every phase starts on a bucket boundary)

    def step_device(_i=None):
        i = pos[0]
        pos[0] += 1
        if world > 1:
            k, j, base = shard.bucket_slot(i, KB, NB)
            if j == 0:
                # NVLink broadcast of the next KB blocks from the ingest GPU; the broadcast of bucket b+1 overlaps the
                # kernels of bucket b
                cur = torch.cuda.current_stream()
                if consumed[k] is not None:
                    cur.wait_event(consumed[k])
                dist.broadcast(d_blocks[base:base + KB] if rank == 0 else d_stage[k], src=0)
                ev = torch.cuda.Event()
                ev.record(cur)
                st.wait_event(ev)
            blk = d_blocks[base + j] if rank == 0 else d_stage[k][j]
            fe.submit_device(cuda.FMT_CF32, blk.data_ptr(), BLOCK)
            if j == KB - 1:
                consumed[k] = torch.cuda.Event()
                consumed[k].record(st)
        else:
            fe.submit_device(cuda.FMT_CF32, d_blocks[i % NB].data_ptr(), BLOCK)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value) ------------------------------------------------------------
    fe.set_readback(False)
    for i in range(args.warmup):
        step_device(i)
    barrier()
    pos[0] = 0
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = fe.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(st)
    for i in range(args.steps):
        step_device(args.warmup + i)
    e1.record(st)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = fe.launches - l0
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    value = args.steps * BLOCK / (ms * 1e-3) / 1e6

    # ---- per-kernel-family device time (CUDA events on the front end's stream, inside the library) ----
    fe.set_profiling(True)
    fam_steps = []
    nprof = max(4, min(args.steps, 16))
    barrier()
    pos[0] = 0
    consumed[0] = consumed[1] = None
    for i in range(nprof):
        step_device(i)
        fe.wait()
        fam_steps.append(fe.kernel_ms())
    # mean over the profiled steps without outliers: a bracket that happens to span a host hiccup (the launches of a
    # family are enqueued one by one between its two events) would otherwise leak into the roofline figures. The plain
    # mean is kept otherwise because the spectrum family is empty in the steps that complete no frame.
    fs = np.array(fam_steps)
    fam = np.zeros(fs.shape[1])
    for c in range(fs.shape[1]):
        col = fs[:, c]
        pos = col[col > 0]
        keep = col[col <= 4.0 * np.median(pos)] if len(pos) else col
        fam[c] = float(np.mean(keep)) if len(keep) else 0.0
    fe.set_profiling(False)

    # ---- end to end through the host API (e2e): pinned host block -> H2D -> path -> D2H results --------
    fe.set_readback(True)
    e2e = None
    if world == 1:
        pin = [cuda.PinnedArray((BLOCK,), np.complex64) for _ in range(4)]
        for j, p in enumerate(pin):
            p.array[:] = host[j % NB]
        touched = 0.0

        def step_host(i):
            fe.submit(cuda.FMT_CF32, pin[i % len(pin)], BLOCK)

        def consume():
            nonlocal touched
            fe.wait()
            iq, dm = fe.vfo_output(ids[0], copy=False)
            rows = fe.fft_rows(copy=False)
            touched += float(iq[0].real) if len(iq) else 0.0
            touched += float(rows[0, 0]) if len(rows) else 0.0

        for i in range(args.warmup):
            step_host(i); consume()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ahead = max(1, min(args.e2e_ahead, 4, args.steps))
        for i in range(ahead):
            step_host(i)
        for i in range(ahead, args.steps):
            step_host(i)      # block i is copied in while blocks i-ahead .. i-1 are still in flight (five result sets)
            consume()         # results of block i-ahead
        for i in range(ahead):
            consume()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        d2h = 0
        for vid, v in zip(ids, mine):
            n = len(fe.vfo_output(vid, copy=False)[0])
            d2h += n * (8 + 4)
        d2h += int(FFT_N * 4 * BLOCK / FFT_N)
        # the host link on its own: the same pinned block copied host -> device back to back (what bounds e2e for cf32 input)
        hsrc = torch.from_numpy(pin[0].array.view(np.float32))
        hdst = torch.empty_like(d_blocks[0].view(-1))
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            hdst.copy_(hsrc, non_blocking=True)
        torch.cuda.synchronize()
        c0.record()
        for _ in range(20):
            hdst.copy_(hsrc, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_gbs = 20 * BLOCK * 8 / (c0.elapsed_time(c1) * 1e-3) / 1e9
        e2e = {"value": args.steps * BLOCK / dt / 1e6, "unit": "MS/s", "h2d_bytes_per_step": BLOCK * 8, "d2h_bytes_per_step": int(d2h),
               "h2d_link_gbs": h2d_gbs, "h2d_link_bound_msps": h2d_gbs * 1e9 / 8 / 1e6,
               "note": "pinned host block -> cudaMemcpyAsync H2D -> full path -> D2H of all VFO outputs + spectrum rows; wall clock. "
                       "h2d_link_gbs: the same 4.9 MB pinned block copied back to back with nothing else running = the ceiling of any cf32 e2e number on this host link"}
        for p in pin:
            p.free()
    else:
        # multi-GPU: the host edge is rank 0's; report the same sharded run fed from pinned host memory on rank 0
        pin = cuda.PinnedArray((BLOCK,), np.complex64) if rank == 0 else None
        if rank == 0:
            pin.array[:] = host[0]
            pin_t = torch.from_numpy(pin.array.view(np.float32).reshape(BLOCK, 2))

        barrier()
        consumed[0] = consumed[1] = None

        nslots = 2 * KB           # the staging buckets of the device-resident leg, used block by block here
        cons_e = [None] * nslots

        def step_e2e(i):
            sl = i % nslots
            buf = d_stage[sl // KB][sl % KB]   # end to end keeps one broadcast per block: each block comes from the host as it is submitted
            cur = torch.cuda.current_stream()
            if cons_e[sl] is not None:
                cur.wait_event(cons_e[sl])
            if rank == 0:
                buf.copy_(pin_t, non_blocking=True)   # H2D from pinned host memory, every step
            dist.broadcast(buf, src=0)
            ev = torch.cuda.Event(); ev.record(cur); st.wait_event(ev)
            fe.submit_device(cuda.FMT_CF32, buf.data_ptr(), BLOCK)
            cons_e[sl] = torch.cuda.Event(); cons_e[sl].record(st)

        sink = 0.0

        def consume():
            nonlocal sink
            fe.wait()                                  # results of the oldest outstanding block are on the host
            iq, _ = fe.vfo_output(ids[0], copy=False)
            sink += float(iq[0].real) if len(iq) else 0.0

        for i in range(args.warmup):
            step_e2e(i); consume()
        barrier()
        ahead = max(1, min(args.e2e_ahead, 4, nslots - 1, args.steps))
        t0 = time.perf_counter()
        for i in range(ahead):
            step_e2e(i)
        for i in range(ahead, args.steps):
            step_e2e(i)       # blocks i-ahead .. i-1 are still in flight (five result sets)
            consume()
        for i in range(ahead):
            consume()
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        d2h = sum(len(fe.vfo_output(vid, copy=False)[0]) * 12 for vid in ids)
        dd = torch.tensor([d2h], dtype=torch.int64, device=dev)
        dist.all_reduce(dd, op=dist.ReduceOp.SUM)
        e2e = {"value": args.steps * BLOCK / float(t.item()) / 1e6, "unit": "MS/s", "h2d_bytes_per_step": BLOCK * 8,
               "d2h_bytes_per_step": int(dd.item()) + int(FFT_N * 4 * BLOCK / FFT_N),
               "note": "rank 0 pinned host block -> H2D -> NCCL broadcast -> sharded path -> per-rank D2H; wall clock, max over ranks"}

    # ---- the spectrum kernels with a full machine's worth of frames (rank 0, N = 1) --------------------------
    # One streaming step completes at most one 1M-point frame = 128 CTAs; this leg shows what the same kernels
    # sustain when every SM has work: FB frames back to back from HBM (FB x 8 MB in, FB x 4 MB out, > L2).
    spec_batched = None
    if rank == 0 and world == 1:
        FB = min(24, d_blocks.numel() // (2 * FFT_N))
        src = d_blocks.view(-1)[: 2 * FB * FFT_N].view(FB * FFT_N, 2) if d_blocks.numel() >= 2 * FB * FFT_N else None
        if src is not None and FB >= 4:
            rows_dev = torch.empty((FB, FFT_N), dtype=torch.float32, device=dev)
            win = cuda.design_window(cuda.WIN_BH4, FFT_N)
            sptr = st.cuda_stream
            for _ in range(3):
                cuda.spectrum_device(FFT_N, FFT_N, FB, FFT_N, src.data_ptr(), win, rows_dev.data_ptr(), sptr)
            torch.cuda.synchronize()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            b0.record(st)
            for _ in range(reps):
                cuda.spectrum_device(FFT_N, FFT_N, FB, FFT_N, src.data_ptr(), None, rows_dev.data_ptr(), sptr)  # cached window
            b1.record(st)
            torch.cuda.synchronize()
            bms = b0.elapsed_time(b1) / reps
            spec_batched = {"frames_per_call": FB, "ms_per_call": bms, "msps": FB * FFT_N / (bms * 1e-3) / 1e6,
                            "achieved": 12.0 * FB * FFT_N / (bms * 1e-3) / 1e9, "unit": "GB/s",
                            "note": "sdrpp_cuda_spectrum_device: FB x 1M-pt frames per call from HBM-resident input (> L2), 12 B/sample algorithmic, CUDA events"}
            del rows_dev

    # ---- roofline + baseline objects (rank 0) -------------------------------------------------------------
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        model = algorithmic_model(mine)
        s1_ms, fft_ms, tail_ms, ingest_ms = fam[2], fam[1], fam[3], fam[0]
        # per launch: the block read once (8 B/sample) + that group's outputs; a step has one launch per VFO class
        n_s1_launches = len({(v[0], v[1], v[3]) for v in mine})
        chan_bytes = (8.0 * n_s1_launches + (model["chan_bytes_per_sample"] - 8.0)) * BLOCK
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["kernels"]
            want = "s1t_" if fe.stage1_tensor_launches > 0 else "stage1_kernel"
            traffic = sum(k["dram_read_bytes"] + k["dram_write_bytes"] for name, k in tj.items() if want in name) or None
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": "stage1_kernel (NCO folded into the first decimating FIR; both VFO-class launches of a step)",
                "achieved": chan_bytes / (s1_ms * 1e-3) / 1e9 if s1_ms > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                "frac": (chan_bytes / (s1_ms * 1e-3) / 1e9 / hbm_peak) if s1_ms > 0 else None, "traffic": traffic,
                "traffic_note": "dram__bytes_read+write summed over the step's stage-1 launches, ncu --set full at N=1 (profiles/r1_traffic.json)",
                "peak_source": peak_src, "ms_per_step": float(s1_ms), "launches_per_step": n_s1_launches,
                "algorithmic_bytes_per_step": chan_bytes,
                "note": "minimal-bytes accounting (8 B/sample in + outputs, SURVEY 8d); this kernel is FP32-FMA-bound, see fp32"}
        sm_clk = (clk or {}).get("sm_mhz") or 1965.0
        fma_peak = 148 * 128 * 2 * sm_clk * 1e6 / 1e12
        s1_flops = model["stage1_flops_per_sample"] * BLOCK
        fp32 = {"achieved": s1_flops / (s1_ms * 1e-3) / 1e12 if s1_ms > 0 else None, "peak": fma_peak, "unit": "TFLOP/s",
                "frac": (s1_flops / (s1_ms * 1e-3) / 1e12 / fma_peak) if s1_ms > 0 else None,
                "peak_source": f"148 SM x 128 FMA lanes x 2 x {sm_clk:.0f} MHz (median SM clock sampled under load)",
                "algorithmic_flops_per_sample": model["stage1_flops_per_sample"]}
        if fe.stage1_tensor_launches > 0:
            # Stage 1 ran on the tensor cores (channelizer_tc.cu). Algorithmic flops stay the reference's own count for
            # NCO + first decimating FIR; the kernel executes more: complex x complex products, three fp16
            # products per fp32 product, tap matrix padded to whole rows, 128-row tiles for 120 outputs.
            tens_peak = float(peaks.get("bf16_tflops", 2250.0))
            tens_src = ("MEASURED_PEAKS.json bf16_tflops (cuBLAS dense, burst; fp16 runs at the same rate)" if "bf16_tflops" in peaks
                        else "fallback 2250 TFLOP/s nominal dense fp16 (B200_PROFILING.md)")
            executed = 0.0
            for (osr, bw, dm), cnt in collections.Counter((v[0], v[1], v[3]) for v in mine).items():
                plan = cuda.design_resampler(SR, osr)[0]
                st1 = cuda.design_decim_plan(plan["predec"])[0] if plan["predec"] > 1 else None
                if st1 is None or st1[0] not in (32, 64):
                    continue
                D, T = int(st1[0]), len(st1[1])
                A = -(-T // D)
                executed += (-(-cnt // 16)) * (-(-(BLOCK // D) // 120)) * 3 * (2 * D // 16) * 2.0 * 128 * (32 * A) * 16
            hbm_view = {k: roof[k] for k in ("achieved", "peak", "unit", "frac", "algorithmic_bytes_per_step", "note")}
            hbm_view["note"] = "minimal-bytes accounting (8 B/sample in + outputs, SURVEY 8d)"
            roof = {"bound": "tensor",
                    "kernel": "s1t_split_kernel + s1t_kernel (tcgen05: NCO + first decimating FIR of every VFO as one split-fp16 matrix product per step)",
                    "achieved": fp32["achieved"], "peak": tens_peak, "unit": "TFLOP/s",
                    "frac": fp32["achieved"] / tens_peak if fp32["achieved"] else None,
                    "traffic": traffic, "traffic_note": roof["traffic_note"], "peak_source": tens_src,
                    "ms_per_step": float(s1_ms), "launches_per_step": 2,
                    "algorithmic_flops_per_step": s1_flops, "algorithmic_flops_per_sample": model["stage1_flops_per_sample"],
                    "executed": {"tflops": executed / (s1_ms * 1e-3) / 1e12 if s1_ms > 0 else None,
                                 "frac_of_peak": executed / (s1_ms * 1e-3) / 1e12 / tens_peak if s1_ms > 0 else None,
                                 "flops_per_step": executed,
                                 "note": "tensor-core flops issued: 3 fp16 products x complex x (tap matrix padded to ceil(T/D) rows) x 128-row tiles per 120 outputs"},
                    "hbm": hbm_view, "fp32_equivalent": fp32,
                    "note": "algorithmic flops = the reference's count for NCO + first FIR (what the FP32 kernel of mode 1 executes); ms_per_step covers the fp16 split of the block and the matrix-product kernel"}
        else:
            roof["fp32"] = fp32
        fft_bytes = 12.0 * BLOCK
        roof_fft = {"bound": "hbm", "kernel": "fft_cols_kernel + fft_rows_kernel (window + 1M-pt FFT + dB row)",
                    "achieved": fft_bytes / (fft_ms * 1e-3) / 1e9 if fft_ms > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                    "frac": (fft_bytes / (fft_ms * 1e-3) / 1e9 / hbm_peak) if fft_ms > 0 else None, "traffic": None,
                    "ms_per_step": float(fft_ms), "algorithmic_bytes_per_step": fft_bytes,
                    "msps": BLOCK / (fft_ms * 1e-3) / 1e6 if fft_ms > 0 else None}
        if spec_batched:
            spec_batched["peak"] = hbm_peak
            spec_batched["frac"] = spec_batched["achieved"] / hbm_peak
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, info = cpu_reference(nblocks=2)
            cpu = {"value": v, "unit": "MS/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"],
                   "channelizer_msps": info["channelizer_msps"], "spectrum_msps": info["spectrum_msps"]}
        line = {
            "metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "block": BLOCK, "vfos": NVFO, "fft": FFT_N, "vfos_per_gpu": len(mine),
                       "l2": f"inputs cycle through {NB} distinct blocks = {NB * BLOCK * 8 / 1e6:.0f} MB (> 126 MB L2)",
                       "parallelism": f"vfo-shard x{world} + NCCL broadcast of {KB}-block buckets" if world > 1 else "1 GPU"},
            "clocks": clk, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "roofline_spectrum": roof_fft, "spectrum_batched": spec_batched,
            "kernel_ms_per_step": {"ingest": float(ingest_ms), "spectrum": float(fft_ms), "channelizer_stage1": float(s1_ms), "channelizer_tail": float(tail_ms)},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    fe.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--input-blocks", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-ahead", type=int, default=4, help="blocks submitted ahead of the one being consumed in the end-to-end leg (1..4)")
    ap.add_argument("--bcast-blocks", type=int, default=4, help="N > 1: IQ blocks per NCCL broadcast (bucket size)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
