"""The BASELINE.json workloads as data: stream rate and format, block size, front-end decimation,
spectrum size / window (saturated mode: every sample enters exactly one frame) and the VFO set. Shared by bench.py
(--config N) and the whole-config parity tests; no oracle import here.

  2  rtl_sdr-style uint8 IQ at 3.2 MS/s, blocks of 7,936: conversion + 131072-pt Hann spectrum + 8 NFM VFOs
  3  20 MS/s cf32, blocks of 100,000: 1M-pt spectrum + 100 WFM VFOs (200 k -> 250 kS/s, quadrature)
  4  61.44 MS/s int16, blocks of 307,200: PowerDecimator x4 front end + 1M-pt Blackman-Harris-4 + 256 USB / AM VFOs
  5  122.88 MS/s cf32, blocks of 614,400: 1M-pt Blackman-Harris-4 + 512 NFM / AM VFOs (the metric's configuration)
"""
import numpy as np

from . import synth

FMT_CF32, FMT_U8_RTL, FMT_I16_FILE = 0, 1, 4
WIN_HANN, WIN_BH4, WIN_BH7 = 2, 5, 6
DEMOD_QUAD, DEMOD_AM, DEMOD_USB = 1, 2, 3

NFM = (48e3, 12.5e3, DEMOD_QUAD)   # outSR, bw, demod: decoder_modules/radio/src/demodulators/nfm.h:65-67
AM = (24e3, 12e3, DEMOD_AM)        # am.h:112-114
USB = (48e3, 2.7e3, DEMOD_USB)     # usb.h:103-105
WFM = (250e3, 200e3, DEMOD_QUAD)   # north-star figure (this fork's radio uses 240 k, wfm.h:246)


class Workload:
    def __init__(self, idx, name, sr, block, fmt, decim, fft_size, fft_window, classes, nvfo, seed):
        self.idx, self.name, self.sr, self.block, self.fmt, self.decim = idx, name, sr, block, fmt, decim
        self.fft_size, self.fft_window, self.seed = fft_size, fft_window, seed
        self.eff_sr = sr / decim
        self.fft_rate = self.eff_sr / fft_size          # saturated: interval = N, skip = 0
        offs = synth.vfo_grid(nvfo, self.eff_sr)
        self.vfos = [(classes[i % len(classes)][0], classes[i % len(classes)][1], float(offs[i]), classes[i % len(classes)][2])
                     for i in range(nvfo)]
        self.bytes_per_sample = {FMT_CF32: 8, FMT_U8_RTL: 2, FMT_I16_FILE: 4}[fmt]
        self.np_dtype = {FMT_CF32: np.complex64, FMT_U8_RTL: np.uint8, FMT_I16_FILE: np.int16}[fmt]
        self.scalars_per_sample = 1 if fmt == FMT_CF32 else 2

    @property
    def nvfo(self):
        return len(self.vfos)

    def tone_vfos(self):
        """Indices of the VFOs that get a carrier in the synthetic stream: four VFOs of each class spread over the set
        plus the very last ones, so spot checks find channels with content in both classes."""
        n = self.nvfo
        step = max(1, n // 4)
        idx = sorted(set(list(range(0, n, step)) + list(range(1, n, step)) + [n - 2, n - 1]) & set(range(n)))
        return idx

    def make_blocks(self, nblocks, seed=None):
        """Synthetic IQ of the workload: carriers (unmodulated, -26 dBFS each) at tone_vfos() centres + white noise at
        -40 dBFS, float32 arithmetic, then quantised to the stream's sample format. Returns an array of shape
        (nblocks, block * scalars_per_sample) in the raw dtype."""
        rng = np.random.Generator(np.random.PCG64(self.seed if seed is None else seed))
        n = nblocks * self.block
        out = np.empty(n, dtype=np.complex64)
        offs = [self.vfos[i][2] for i in self.tone_vfos()]
        # carriers sit a little off the VFO centre so that FM / SSB demods see a beat, not DC
        offs = [o + 700.0 for o in offs]
        chunk = 1 << 20
        sigma = np.float32(10.0 ** (-40.0 / 20.0) / np.sqrt(2.0))
        amp = np.float32(min(0.05, 0.6 / max(1, len(offs))))
        for s in range(0, n, chunk):
            m = min(chunk, n - s)
            t = np.arange(s, s + m, dtype=np.float64) / self.sr
            x = np.zeros(m, dtype=np.complex64)
            for f in offs:
                ph = (2.0 * np.pi) * ((f * t) % 1.0)
                x += amp * (np.cos(ph) + 1j * np.sin(ph)).astype(np.complex64)
            x += sigma * (rng.standard_normal(m, dtype=np.float32) + 1j * rng.standard_normal(m, dtype=np.float32))
            out[s:s + m] = x
        if self.fmt == FMT_CF32:
            return out.reshape(nblocks, self.block)
        raw = synth.quantise(out.astype(np.complex128), self.fmt)
        return raw.reshape(nblocks, self.block * 2)

    def describe(self):
        fmt = {FMT_CF32: "cf32", FMT_U8_RTL: "uint8 (rtl formula)", FMT_I16_FILE: "int16 (file formula)"}[self.fmt]
        win = {WIN_HANN: "Hann", WIN_BH4: "Blackman-Harris-4", WIN_BH7: "Blackman-Harris-7"}[self.fft_window]
        fe = f"; PowerDecimator x{self.decim} front end -> {self.eff_sr / 1e6:g} MS/s" if self.decim > 1 else ""
        return (f"{self.sr / 1e6:g} MS/s {fmt} IQ, blocks of {self.block}{fe}; saturated {self.fft_size}-pt {win} spectrum; {self.name}")


def config(idx):
    """BASELINE.json configs[idx-1] (idx 2..5)."""
    if idx == 2:
        return Workload(2, "8 NFM VFOs 12.5k->48k (quadrature)", 3.2e6, 7936, FMT_U8_RTL, 1, 131072, WIN_HANN, [NFM], 8, 2)
    if idx == 3:
        return Workload(3, "100 WFM VFOs 200k->250k (quadrature)", 20e6, 100000, FMT_CF32, 1, 1 << 20, WIN_BH7, [WFM], 100, 3)
    if idx == 4:
        return Workload(4, "256 VFOs alternating USB 2.7k->48k (xlate + real) / AM 12k->24k (magnitude)", 61.44e6, 307200,
                        FMT_I16_FILE, 4, 1 << 20, WIN_BH4, [USB, AM], 256, 4)
    if idx == 5:
        return Workload(5, "512 VFOs alternating NFM 12.5k->48k (quadrature) / AM 12k->24k (magnitude)", 122.88e6, 614400,
                        FMT_CF32, 1, 1 << 20, WIN_BH4, [NFM, AM], 512, 5)
    raise ValueError("config must be 2, 3, 4 or 5")
