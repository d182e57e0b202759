#!/usr/bin/env python3
"""bench.py -- sustained input MS/s of the SDR++ signal-path hot loop on B200.

Default workload = BASELINE.json configs[4] + the 1M-point spectrum the metric is quoted on (--config 5): 122.88 MS/s
complex64 IQ in blocks of 614,400 samples (sr/200), a saturated 1,048,576-point Blackman-Harris-4 spectrum (every sample
enters one frame) and 512 VFOs alternating NFM (12.5 kHz -> 48 kS/s, quadrature demod) and AM (12 kHz -> 24 kS/s,
magnitude). --config 2 / 3 / 4 time the other BASELINE configurations (sdrpp_b200/workloads.py) with their packed sample
formats going over the host link packed. A "step" is one IQ block through conversion/ingest -> [front-end decimation] ->
spectrum frames -> channelizer -> demod front ends.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--impl reference]

N > 1 (torchrun, one rank per GPU): the VFO set is sharded across ranks; the library itself (sdrpp_cuda_comm_*, NCCL
over NVLink) broadcasts every raw block from rank 0 inside sdrpp_cuda_frontend_submit*, the spectrum stays on rank 0
(SURVEY 8e). torch.distributed is used for the barrier and the max-over-ranks reductions only.
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

from sdrpp_b200 import workloads  # noqa: E402

METRIC = "sustained input MS/s (1M-pt FFT + N-VFO channelizer); % of B200 HBM roofline"


# ---------------------------------------------------------------------------------------------
# algorithmic work per input sample
# ---------------------------------------------------------------------------------------------
def algorithmic_model(w, vfos):
    """Per RAW input sample of workload w, for the VFO list `vfos` (a rank's shard): the reference's own flop count
    (NCO 8 + 4 per tap and output for every FIR) and the minimal bytes of SURVEY 8d."""
    from sdrpp_b200 import cuda
    flops = s1_flops = out_bytes = 0.0
    for (osr, bw, _off, demod) in vfos:
        info, _ = cuda.design_resampler(w.eff_sr, osr)
        f, rate = 8.0, 1.0
        stages = cuda.design_decim_plan(info["predec"]) if info["mode"] in (0, 1) else []
        for i, (d, taps) in enumerate(stages):
            c = 4.0 * len(taps) * rate / d
            f += c
            if i == 0:
                s1_flops += 8.0 + c
            rate /= d
        if not stages:
            s1_flops += 8.0
        if info["mode"] in (0, 2):
            f += 4.0 * info["tpp"] * (osr / w.eff_sr)
        if bw != osr:
            f += 4.0 * int(3.8 * osr / (bw / 20.0)) * (osr / w.eff_sr)
        flops += f
        out_bytes += (osr / w.eff_sr) * (8 + (4 if demod else 0))
    k = 1.0 / w.decim   # per raw sample
    fe_flops = 0.0
    if w.decim > 1:
        rate = 1.0
        for d, taps in cuda.design_decim_plan(w.decim):
            fe_flops += 4.0 * len(taps) * rate / d
            rate /= d
    return dict(flops_per_sample=flops * k, stage1_flops_per_sample=s1_flops * k, chan_bytes_per_sample=(8.0 + out_bytes) * k,
                convert_bytes_per_sample=w.bytes_per_sample + 8.0, decim_bytes_per_sample=(8.0 + 8.0 / w.decim) if w.decim > 1 else 0.0,
                decim_flops_per_sample=fe_flops, fft_bytes_per_sample=(w.bytes_per_sample + 4.0) * k if w.decim == 1 else 12.0 * k,
                fft_flops_per_sample=(5.0 * np.log2(w.fft_size) + 12) * k)


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


def cpu_reference(w, nblocks=2, sample_vfos=None, fft_frames=1):
    """Times the reference's CPU implementation of workload w on a bounded sample: `sample_vfos` of its VFOs
    (thread-per-VFO multiplexed on all host cores, like the reference's thread-per-block model) on nblocks blocks, one
    spectrum line on one thread, and -- where the workload has them -- the source conversion loop and the front-end
    PowerDecimator on one thread each (the reference runs every block on a thread of its own: the stream advances at
    the pace of the slowest stage). Returns (MS/s of raw input, info dict)."""
    from oracle import pyoracle as po
    cores = os.cpu_count() or 1
    lib = po.Ref("fast") if po.have_ref("fast") else (po.Ref("") if po.have_ref("") else None)
    port = po.Port()
    kind = "reference" if lib is not None else "port"
    vf = w.vfos
    if sample_vfos is None:
        sample_vfos = min(w.nvfo, 2 * cores)
    pick = [vf[(i * w.nvfo) // sample_vfos] for i in range(sample_vfos)]
    need = max(w.block, w.fft_size * w.decim)
    nb = -(-need // w.block)
    key = ("raw", w.idx)
    if key not in _CPU_CACHE:
        _CPU_CACHE[key] = w.make_blocks(nb)
    raw = _CPU_CACHE[key]
    stages = {}
    # source conversion (A1) on one thread
    if w.fmt != po.FMT_CF32:
        t0 = time.perf_counter()
        x_blocks = [port.convert(w.fmt, raw[b]) for b in range(nb)]
        stages["convert"] = (time.perf_counter() - t0) / (nb * w.block)
    else:
        x_blocks = [raw[b] for b in range(nb)]
    # front-end PowerDecimator (A3) on one thread
    if w.decim > 1:
        pd = (lib or port).powerdecim(w.decim)
        t0 = time.perf_counter()
        x_blocks = [pd.process(x) for x in x_blocks]
        stages["front_end_decimation"] = (time.perf_counter() - t0) / (nb * w.block)
    stream = np.ascontiguousarray(np.concatenate(x_blocks))
    blk_eff = w.block // w.decim
    if lib is not None:
        win = lib.window(w.fft_window, w.fft_size)
        t_ch = lib.bench_channelizer(w.eff_sr, pick, stream[:blk_eff], nblocks, min(cores, sample_vfos))
        t_fft = lib.bench_channelizer(w.eff_sr, pick[:1], stream[:w.fft_size], 0, 1, fft=(w.fft_size, win, fft_frames))
        threads = min(cores, sample_vfos)
    else:
        t0 = time.perf_counter()
        for v in pick:
            o = port.rxvfo(w.eff_sr, v[0], v[1], v[2])
            for _ in range(nblocks):
                o.process(stream[:blk_eff])
        t_ch = time.perf_counter() - t0
        t0 = time.perf_counter()
        win = port.window(w.fft_window, w.fft_size)
        for _ in range(fft_frames):
            port.spectrum(w.fft_size, stream[:w.fft_size], win, want64=False)
        t_fft = time.perf_counter() - t0
        threads = 1
    stages["channelizer"] = (t_ch * (w.nvfo / float(sample_vfos))) / (nblocks * w.block)
    stages["spectrum"] = t_fft / (fft_frames * w.fft_size * w.decim)
    slowest = max(stages, key=stages.get)
    msps = 1e-6 / stages[slowest]
    info = dict(kind=kind, cores=threads,
                sample=f"{sample_vfos} of {w.nvfo} VFOs x {nblocks} blocks of {w.block} (scaled x{w.nvfo / sample_vfos:g}) on {threads} threads "
                       f"+ {fft_frames} x {w.fft_size}-pt spectrum line on 1 thread"
                       + (" + source conversion loop on 1 thread" if "convert" in stages else "")
                       + (f" + PowerDecimator x{w.decim} on 1 thread" if w.decim > 1 else "")
                       + "; reference dsp/ headers, generic-VOLK shim, -O3 -ffast-math; stream rate = the slowest stage",
                stage_msps={k: 1e-6 / v for k, v in stages.items()}, slowest_stage=slowest, seconds=t_ch + t_fft)
    return msps, info


def config_dict(w, **extra):
    d = {"workload": w.describe(), "baseline_config": w.idx, "block": w.block, "vfos": w.nvfo, "fft": w.fft_size}
    d.update(extra)
    return d


def run_reference(args, rank, world):
    if rank != 0:
        return
    w = workloads.config(args.config)

    vals, info = [], None
    for i in range(args.warmup + args.steps):
        v, info = cpu_reference(w, nblocks=1)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * w.block / (value * 1e6), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(w),
        "cpu_baseline": {"value": value, "unit": "MS/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"],
                         "stage_msps": info["stage_msps"], "slowest_stage": info["slowest_stage"]},
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
    from sdrpp_b200 import cuda, shard

    if cuda.device_count() <= 0:
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    w = workloads.config(args.config)
    # ---- the C++ interface modules link against (IQFrontEnd + VFOManager + dsp::stream), N = 1 -------------------------
    # First of all, before this process has a CUDA context of its own: the demo is a separate process, and beside a parent that
    # holds a context, 160 MB of source blocks and its own threads it measured 2.7-3.0 GS/s instead of the 3.9-4.4 of a
    # stand-alone run.
    e2e_cpp = None
    demo = os.path.join(ROOT, "tests", "cpp", "mirror_demo")
    if rank == 0 and world == 1 and w.fmt == workloads.FMT_CF32 and os.path.exists(demo) and not args.no_cpp:
        try:
            r = subprocess.run([demo, "bench", repr(w.sr), str(w.block), str(w.fft_size), str(w.nvfo), str(max(300, min(600, args.steps))), "8"],
                               capture_output=True, text=True, timeout=300)
            last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            if r.returncode == 0 and last:
                e2e_cpp = json.loads(last[-1])
                e2e_cpp["note"] = ("tests/cpp/mirror_demo bench: a source thread swaps pinned blocks into the input dsp::stream, sigpath::iqFrontEnd "
                                   "(ingest / deliver / spectrum threads over the C ABI) feeds VFOs created through sigpath::vfoManager, 8 consumer "
                                   "threads read()/flush() every VFO::output stream, acquire/release receive every spectrum row")
            else:
                e2e_cpp = {"error": (r.stderr or r.stdout)[-300:]}
        except Exception as ex:  # noqa: BLE001
            e2e_cpp = {"error": str(ex)}
    torch.cuda.set_device(local_rank)
    cuda.init(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = None
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: NCCL's own log (NCCL_DEBUG, if the caller set it) goes to stderr
        if os.environ.get("NCCL_DEBUG") and not os.environ.get("NCCL_DEBUG_FILE"):
            os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
        dist.init_process_group("nccl", device_id=dev)
        # the library owns the communicator of the data path; torch.distributed only carries its 128-byte id
        box = [cuda.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm = cuda.Comm(box[0], rank, world, local_rank)

    # shard the VFO set across ranks, balanced by per-VFO cost (sdrpp_b200/shard.py); rank 0 also ingests, broadcasts and
    # runs the spectrum, so it gets a base load
    costs = [shard.vfo_cost(w.eff_sr, v[0], v[1], cuda.design_resampler, cuda.design_decim_plan) for v in w.vfos]
    base = [args.rank0_base_load] + [0.0] * (world - 1) if world > 1 else None
    mine_idx = shard.shard_vfos(costs, world, base)[rank]
    mine = [w.vfos[i] for i in mine_idx]
    with_fft = (rank == 0)

    def new_frontend():
        f = cuda.Frontend(w.sr, decim_ratio=w.decim, fft_size=w.fft_size if with_fft else 0, fft_rate=w.fft_rate,
                          fft_window=w.fft_window, max_block=w.block)
        if comm is not None:
            f.set_comm(comm, 0)
        return f

    fe = new_frontend()
    ids = [fe.add_vfo(*v) for v in mine]
    st = torch.cuda.ExternalStream(fe.stream, device=dev)

    # ---- source blocks resident in HBM on the ingest rank, laid out over more than the L2 (126 MB) --------------------
    blk_bytes = w.block * w.bytes_per_sample
    nbase = max(4, min(args.input_blocks, -(-(160 << 20) // blk_bytes)))
    NB = max(nbase, -(-(160 << 20) // blk_bytes)) if rank == 0 else 0
    host = w.make_blocks(nbase) if rank == 0 else None
    d_src = None
    if rank == 0:
        hb = torch.from_numpy(host.view(np.uint8).reshape(nbase, blk_bytes))
        d_src = torch.empty((NB, blk_bytes), dtype=torch.uint8, device=dev)
        for s in range(0, NB, nbase):
            n = min(nbase, NB - s)
            d_src[s:s + n].copy_(hb[:n])
    torch.cuda.synchronize()

    def step_device(i):
        if rank == 0:
            fe.submit_device(w.fmt, d_src[i % NB].data_ptr(), w.block)   # N > 1: the library broadcasts it from here
        else:
            fe.submit_shared(w.fmt, w.block)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum_int(x):
        if world == 1:
            return int(x)
        t = torch.tensor([x], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return int(t.item())

    # ---- device-resident timing (value) ------------------------------------------------------------
    fe.set_readback(False)
    # Untimed: the front end instantiates a CUDA graph for a block's command sequence the second time it sees it (~30
    # distinct sequences in a steady stream: result slot x region parity x frame-completing or not). A continuous stream is
    # past this after a fraction of a second; the benchmark runs it before its warm-up steps so that the timed steps are
    # steady-state ones.
    for i in range(max(0, args.graph_warmup)):
        step_device(i)
    barrier()
    for i in range(args.warmup):
        step_device(i)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = fe.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(st)
    for i in range(args.steps):
        step_device(args.warmup + i)
    fe.join_streams()      # the event below covers the spectrum, tail and broadcast streams of the last blocks too
    e1.record(st)
    barrier()
    ms = reduce_max(e0.elapsed_time(e1))
    launches = reduce_sum_int(fe.launches - l0)
    clk = clocks.stop() if rank == 0 else None
    value = args.steps * w.block / (ms * 1e-3) / 1e6
    tensor_launches = reduce_sum_int(fe.stage1_tensor_launches)

    # ---- per-kernel-family device time (CUDA events on the front end's stream, inside the library) ----
    fe.set_profiling(True)
    fam_steps = []
    nprof = max(4, min(args.steps, 16))
    barrier()
    for i in range(nprof):
        step_device(i)
        fe.wait()
        fam_steps.append(fe.kernel_ms())
    # mean over the profiled steps without outliers (a bracket that spans a host hiccup); plain mean otherwise because
    # the spectrum family is empty in the steps that complete no frame
    fs = np.array(fam_steps)
    fam = np.zeros(fs.shape[1])
    for c in range(fs.shape[1]):
        col = fs[:, c]
        nz = col[col > 0]
        keep = col[col <= 4.0 * np.median(nz)] if len(nz) else col
        fam[c] = float(np.mean(keep)) if len(keep) else 0.0
    fe.set_profiling(False)
    barrier()

    # ---- end to end through the host API (e2e): pinned host block -> H2D -> [broadcast] -> path -> D2H results --------
    fe.set_readback(True)
    raw_dtype = w.np_dtype
    pin = None
    if rank == 0:
        pin = [cuda.PinnedArray((w.block * w.scalars_per_sample,), raw_dtype) for _ in range(4)]
        for j, p in enumerate(pin):
            p.array[:] = host[j % nbase]
    sink = [0.0]

    def step_host(i):
        if rank == 0:
            fe.submit(w.fmt, pin[i % len(pin)], w.block)
        else:
            fe.submit_shared(w.fmt, w.block)

    def consume():
        fe.wait()
        if ids:
            iq, _dm = fe.vfo_output(ids[0], copy=False)
            sink[0] += float(iq[0].real) if len(iq) else 0.0
        if with_fft:
            rows = fe.fft_rows(copy=False)
            sink[0] += float(rows[0, 0]) if len(rows) else 0.0

    # untimed, like in front of the device-resident leg: blocks that arrive over the host link are planned differently
    # (their own command sequences), so their graphs are instantiated here and not inside the timed region
    for i in range(max(0, args.graph_warmup)):
        step_host(i); consume()
    for i in range(args.warmup):
        step_host(i); consume()
    barrier()
    graphs_before_e2e = fe.graph_stats()["graphs_instantiated"]
    t0 = time.perf_counter()
    ahead = max(1, min(args.e2e_ahead, 4, args.steps))
    for i in range(ahead):
        step_host(i)
    for i in range(ahead, args.steps):
        step_host(i)      # block i is copied in while blocks i-ahead .. i-1 are still in flight (five result sets)
        consume()         # results of block i-ahead
    for i in range(ahead):
        consume()
    barrier()
    dt = reduce_max(time.perf_counter() - t0)
    d2h = sum(len(fe.vfo_output(vid, copy=False)[0]) * (8 + (4 if v[3] else 0)) for vid, v in zip(ids, mine))
    d2h = reduce_sum_int(d2h) + int(w.fft_size * 4 * (w.block / w.decim) / w.fft_size)
    e2e = {"value": args.steps * w.block / dt / 1e6, "unit": "MS/s", "h2d_bytes_per_step": blk_bytes, "d2h_bytes_per_step": int(d2h),
           "blocks_in_flight": ahead + 1, "graphs_instantiated_inside_timed_region": fe.graph_stats()["graphs_instantiated"] - graphs_before_e2e,
           "note": "pinned host block (packed sample format) -> cudaMemcpyAsync H2D -> " + ("ncclBroadcast inside the library -> " if world > 1 else "")
                   + "full path -> D2H of every VFO's iq + demod rows and every spectrum row; wall clock, max over ranks"}
    if rank == 0:
        # the host link on its own: the same pinned block copied host -> device back to back (the ceiling of any e2e number)
        hsrc = torch.from_numpy(pin[0].array.view(np.uint8))
        hdst = torch.empty(blk_bytes, dtype=torch.uint8, device=dev)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(20, min(2000, int((64 << 20) / blk_bytes)))
        for _ in range(3):
            hdst.copy_(hsrc, non_blocking=True)
        torch.cuda.synchronize()
        c0.record()
        for _ in range(reps):
            hdst.copy_(hsrc, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        h2d_gbs = reps * blk_bytes / (c0.elapsed_time(c1) * 1e-3) / 1e9
        e2e["h2d_link_gbs"] = h2d_gbs
        e2e["h2d_link_bound_msps"] = h2d_gbs * 1e9 / w.bytes_per_sample / 1e6
        # ... and with this rank's result bytes going the other way at the same time, in the proportion of a step: PCIe is
        # full duplex, but the read requests of the block copy share the upstream direction with the result data
        # (tools/e2e_probe.py: 53 -> 41 GB/s with the reverse direction saturated). This is the ceiling of an end-to-end
        # number that reads every result back on this host.
        d2h_rank0 = max(1, int(d2h if world == 1 else sum(len(fe.vfo_output(vid, copy=False)[0]) * (8 + (4 if v[3] else 0)) for vid, v in zip(ids, mine))
                               + int(w.fft_size * 4 * (w.block / w.decim) / w.fft_size)))
        dsrc2 = torch.empty(d2h_rank0, dtype=torch.uint8, device=dev)
        hdst2 = torch.empty(d2h_rank0, dtype=torch.uint8).pin_memory()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        c0.record()
        for _ in range(reps):
            hdst.copy_(hsrc, non_blocking=True)
            with torch.cuda.stream(side):
                hdst2.copy_(dsrc2, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        bidir_gbs = reps * blk_bytes / (c0.elapsed_time(c1) * 1e-3) / 1e9
        e2e["h2d_link_gbs_with_results_going_back"] = bidir_gbs
        e2e["link_bound_bidirectional_msps"] = bidir_gbs * 1e9 / w.bytes_per_sample / 1e6
        e2e["link_note"] = ("h2d_link_bound_msps: the block copy alone, back to back. link_bound_bidirectional_msps: the same copies with this "
                            "rank's result bytes per step issued alternately on a second stream (an estimate, +-5 %, of what the link carries "
                            "when every result is read back: the read requests of the block copy share the upstream direction with the result "
                            "data). The end-to-end value sits at this second figure: it is bound by the host link, not by the kernels.")
        for p in pin:
            p.free()
    barrier()

    # ---- the spectrum kernels with a full machine's worth of frames (rank 0, N = 1) --------------------------
    spec_batched = None
    if rank == 0 and world == 1 and w.fmt == workloads.FMT_CF32 and w.decim == 1:
        N = w.fft_size
        FB = min(24, (NB * w.block) // N)
        if FB >= 4:
            src = d_src.view(-1)[: FB * N * 8]
            rows_dev = torch.empty((FB, N), dtype=torch.float32, device=dev)
            win = cuda.design_window(w.fft_window, N)
            sptr = st.cuda_stream
            for _ in range(3):
                cuda.spectrum_device(N, N, FB, N, src.data_ptr(), win, rows_dev.data_ptr(), sptr)
            torch.cuda.synchronize()
            b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            b0.record(st)
            for _ in range(reps):
                cuda.spectrum_device(N, N, FB, N, src.data_ptr(), None, rows_dev.data_ptr(), sptr)  # cached window
            b1.record(st)
            torch.cuda.synchronize()
            bms = b0.elapsed_time(b1) / reps
            spec_batched = {"frames_per_call": FB, "ms_per_call": bms, "msps": FB * N / (bms * 1e-3) / 1e6,
                            "achieved": 12.0 * FB * N / (bms * 1e-3) / 1e9, "unit": "GB/s",
                            "note": "sdrpp_cuda_spectrum_device: FB frames per call from HBM-resident input (> L2), 12 B/sample algorithmic, CUDA events"}
            del rows_dev
    comm_info = comm.info() if comm is not None else None
    graph_stats = fe.graph_stats()
    fe.close()

    # ---- parity check OUTSIDE the timed region: the same workload, the same feed, against the oracle ------------------
    parity = None
    if not args.no_parity:
        from tests import parity_workload   # test infrastructure (oracle = checker only)

        def feed(f, fmt, raw):
            if rank == 0:
                f.submit(fmt, raw)
            else:
                f.submit_shared(fmt, w.block)

        nb_par = {2: 40, 3: 11, 4: 14, 5: 4}[w.idx]
        res = parity_workload.check_workload(cuda, w, nblocks=nb_par, vfo_ids=mine_idx, with_fft=with_fft, submit=feed if world > 1 else None,
                                             frontend_setup=(lambda f: f.set_comm(comm, 0)) if comm is not None else None,
                                             spot_count=32 if world == 1 else 12)
        ok_all = reduce_sum_int(0 if res["ok"] else 1) == 0
        worst_iq = reduce_max(res["worst_iq_rel_rms"])
        worst_dm = reduce_max(res["worst_demod_stage_isolated"])
        checked = reduce_sum_int(res["vfos_checked"])
        parity = {"ok": ok_all, "ranks": world, "vfos_checked": checked, "gate": res["gate"], "worst_iq_rel_rms_vs_ideal_nco_oracle": worst_iq,
                  "worst_demod_stage_isolated": worst_dm, "counts_exact": res["counts_exact"], "blocks": res["blocks"],
                  "ref_f32_rotator_vs_ideal": res["ref_f32_rotator_vs_ideal"], "adjudicated_rank0": res["adjudicated"],
                  "rows_checked_rank0": res.get("rows"), "worst_row_db_within_100dB_rank0": res.get("worst_row_db_within_100dB"),
                  "failures_rank0": res["failures"],
                  "note": "fresh front end(s) with the benchmarked VFO shards, blocks through the same submit path (N > 1: the library's "
                          "broadcast), spot VFOs of every rank's shard against the ideal-NCO oracle chain; tests/parity_workload.py"}

    # ---- roofline + baseline objects (rank 0) -------------------------------------------------------------
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        model = algorithmic_model(w, mine)
        ingest_ms, fft_ms, s1_ms, tail_ms = fam[0], fam[1], fam[2], fam[3]
        traffic = {}
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))["kernels"]
        except Exception:
            pass

        def traffic_of(pat):
            v = sum(k["dram_read_bytes"] + k["dram_write_bytes"] for name, k in traffic.items() if pat in name)
            return v or None

        def hbm_row(kernel, bytes_per_sample, ms_, note, pat=None):
            by = bytes_per_sample * w.block
            return {"bound": "hbm", "kernel": kernel, "achieved": by / (ms_ * 1e-3) / 1e9 if ms_ > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                    "frac": by / (ms_ * 1e-3) / 1e9 / hbm_peak if ms_ > 0 else None, "traffic": traffic_of(pat) if pat and w.idx == 5 else None,
                    "peak_source": peak_src, "ms_per_step": float(ms_), "algorithmic_bytes_per_step": by, "note": note}

        rows = {}
        # ingest family: conversion (+ front-end decimation where the workload has it)
        pre_bytes = model["convert_bytes_per_sample"] + model["decim_bytes_per_sample"] - (8.0 if w.decim > 1 else 0.0)
        fused_ingest = (w.fmt == workloads.FMT_CF32 and w.decim == 1 and tensor_launches > 0 and os.environ.get("SDRPP_FUSE_INGEST", "1") != "0")
        if fused_ingest:
            # cf32 blocks are not copied to the ring by a kernel of their own: the fp16 split of the tensor-core stage 1 reads the
            # block in place and writes ring + hi/lo planes (8 B in, 8 B ring, 4 + 4 B planes per sample)
            rows["ingest"] = hbm_row("s1t_split_kernel<fused cf32 ingest> (ring write + fp16 hi/lo planes of the tensor-core stage 1)", 24.0, ingest_ms,
                                     "8 B/sample in + 8 B ring + 8 B fp16 hi/lo planes; the separate conversion pass of SURVEY 8d rows 1-2 (16 B/sample) is gone", "s1t_split")
        else:
            rows["ingest"] = hbm_row("ingest_kernel" + (" + decim_stage_kernel x2 (PowerDecimator x%d)" % w.decim if w.decim > 1 else "") + (" + s1t_split_kernel" if tensor_launches > 0 else ""), pre_bytes, ingest_ms,
                                     "conversion %g B/sample in + 8 B out" % w.bytes_per_sample + (" fused with the first decimator stage's input; + 8/%d B out" % w.decim if w.decim > 1 else "")
                                     + " (SURVEY 8d rows 1-2)", "ingest")
        rows["spectrum"] = hbm_row("fft_cols_kernel + fft_rows_kernel (window + %d-pt FFT + dB row)" % w.fft_size, model["fft_bytes_per_sample"], fft_ms,
                                   "SURVEY 8d: sample in (packed size for packed formats) + 4 B out per spectrum sample", "fft_")
        if rows["spectrum"]["ms_per_step"] > 0:
            rows["spectrum"]["msps"] = w.block / (fft_ms * 1e-3) / 1e6
        sm_clk = (clk or {}).get("sm_mhz") or 1965.0
        fma_peak = 148 * 128 * 2 * sm_clk * 1e6 / 1e12
        s1_flops = model["stage1_flops_per_sample"] * w.block
        fp32 = {"achieved": s1_flops / (s1_ms * 1e-3) / 1e12 if s1_ms > 0 else None, "peak": fma_peak, "unit": "TFLOP/s",
                "frac": (s1_flops / (s1_ms * 1e-3) / 1e12 / fma_peak) if s1_ms > 0 else None,
                "peak_source": f"148 SM x 128 FMA lanes x 2 x {sm_clk:.0f} MHz (median SM clock sampled under load)",
                "algorithmic_flops_per_sample": model["stage1_flops_per_sample"]}
        chan_min = hbm_row("channelizer stage 1 (NCO + first decimating FIR of every VFO)", model["chan_bytes_per_sample"], s1_ms,
                           "minimal-bytes accounting of SURVEY 8d: the stream read ONCE (8 B per decimated sample) + every VFO's outputs", "s1t_")
        if tensor_launches > 0:
            tens_peak = float(peaks.get("bf16_tflops", 2250.0))
            tens_src = ("MEASURED_PEAKS.json bf16_tflops (cuBLAS dense, burst; fp16 runs at the same rate)" if "bf16_tflops" in peaks
                        else "fallback 2250 TFLOP/s nominal dense fp16 (B200_PROFILING.md)")
            executed = 0.0
            for (osr, bw, dm), cnt in collections.Counter((v[0], v[1], v[3]) for v in mine).items():
                plan = cuda.design_resampler(w.eff_sr, osr)[0]
                st1 = cuda.design_decim_plan(plan["predec"])[0] if plan["predec"] > 1 else None
                if st1 is None or st1[0] not in (32, 64):
                    continue
                D, T = int(st1[0]), len(st1[1])
                A = -(-T // D)
                executed += (-(-cnt // 16)) * (-(-((w.block // w.decim) // D) // 120)) * 3 * (2 * D // 16) * 2.0 * 128 * (32 * A) * 16
            rows["channelizer_stage1"] = {
                "bound": "tensor",
                "kernel": "s1t_kernel (tcgen05: NCO + first decimating FIR of every VFO as one split-fp16 matrix product per step; the fp16 split of the block is timed with the ingest family)",
                "achieved": fp32["achieved"], "peak": tens_peak, "unit": "TFLOP/s", "frac": fp32["achieved"] / tens_peak if fp32["achieved"] else None,
                "traffic": traffic_of("s1t_kernel") if w.idx == 5 else None,
                "traffic_note": "dram__bytes_read+write of the step's s1t_kernel launch, ncu --set full at N=1 (profiles/r2_traffic.json): 9.5 MB against 9.8 MB of fp16 planes -- the 32 VFO tiles re-read them out of L2",
                "peak_source": tens_src, "ms_per_step": float(s1_ms), "launches_per_step": 1,
                "algorithmic_flops_per_step": s1_flops, "algorithmic_flops_per_sample": model["stage1_flops_per_sample"],
                "executed": {"tflops": executed / (s1_ms * 1e-3) / 1e12 if s1_ms > 0 else None,
                             "frac_of_peak": executed / (s1_ms * 1e-3) / 1e12 / tens_peak if s1_ms > 0 else None, "flops_per_step": executed,
                             "note": "tensor-core flops issued: 3 fp16 products x complex x (tap matrix padded to ceil(T/D) rows) x 128-row tiles per 120 outputs"},
                "hbm": {k: chan_min[k] for k in ("achieved", "peak", "unit", "frac", "algorithmic_bytes_per_step", "note")},
                "fp32_equivalent": fp32,
                "note": "algorithmic flops = the reference's count for NCO + first FIR (SURVEY 8d); ms_per_step is the matrix-product kernel (CUDA events around it inside the library)"}
        else:
            chan_min["kernel"] = "stage1_kernel / mix_only_kernel (FP32: NCO folded into the first decimating FIR, one launch per VFO class)"
            chan_min["fp32"] = fp32
            chan_min["note"] += "; the kernel is FP32-FMA-bound (SURVEY 8d), see fp32"
            rows["channelizer_stage1"] = chan_min
        rows["channelizer_tail"] = {"kernel": "tail_stage0_wide_kernel + tail_kernel (remaining FIR stages, polyphase resampler, channel filter, demod front ends)",
                                    "ms_per_step": float(tail_ms), "note": "latency/barrier-bound (<= 1 FMA per input sample and VFO); runs beside stage 1 of the next block"}
        dominant = max(("ingest", "spectrum", "channelizer_stage1"), key=lambda k: rows[k]["ms_per_step"])
        roof = dict(rows[dominant])
        roof["family"] = dominant
        if spec_batched:
            spec_batched["peak"] = hbm_peak
            spec_batched["frac"] = spec_batched["achieved"] / hbm_peak
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            v, info = cpu_reference(w, nblocks=2)
            cpu = {"value": v, "unit": "MS/s", "cores": info["cores"], "kind": info["kind"], "sample": info["sample"],
                   "stage_msps": info["stage_msps"], "slowest_stage": info["slowest_stage"]}
        dtype = "f32 (stage 1: 3 x fp16-split products, fp32 accumulate)" if tensor_launches > 0 else "f32"
        line = {
            "metric": METRIC, "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic",
            "config": config_dict(w, vfos_per_gpu=len(mine),
                                  l2=f"inputs cycle through {NB} block buffers = {NB * blk_bytes / 1e6:.0f} MB of HBM (> 126 MB L2)",
                                  parallelism=(f"vfo-shard x{world}; one ncclBroadcast of the raw block per step issued by the library (sdrpp_cuda_comm_*), spectrum on rank 0"
                                               if world > 1 else "1 GPU")),
            "clocks": clk, "e2e": e2e, "e2e_cpp": e2e_cpp, "gpu_launches": int(launches), "parity_check": parity,
            "roofline": roof, "rooflines": rows, "roofline_spectrum": rows["spectrum"], "spectrum_batched": spec_batched,
            "kernel_ms_per_step": {"ingest": float(ingest_ms), "spectrum": float(fft_ms), "channelizer_stage1": float(s1_ms), "channelizer_tail": float(tail_ms)},
            "comm": comm_info, "graphs": dict(graph_stats, warmup_blocks=args.graph_warmup,
                                              note="rank 0; stage-1 and tail command runs replayed as instantiated CUDA graphs / executed command by command, whole run"),
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if comm is not None:
        barrier()
        comm.close()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=[2, 3, 4, 5], help="BASELINE.json configuration (default 5: the metric's)")
    ap.add_argument("--input-blocks", type=int, default=32, help="distinct synthetic blocks (tiled over > 126 MB of HBM)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-cpp", action="store_true")
    ap.add_argument("--graph-warmup", type=int, default=192, help="untimed blocks in front of the warm-up steps (CUDA-graph instantiation)")
    ap.add_argument("--e2e-ahead", type=int, default=4, help="blocks submitted ahead of the one being consumed in the end-to-end leg (1..4)")
    ap.add_argument("--rank0-base-load", type=float, default=770.0, help="N > 1: VFO-cost units rank 0 is charged for ingest + spectrum")
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
