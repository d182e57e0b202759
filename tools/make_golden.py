#!/usr/bin/env python3
"""Generate tests/golden/ from the REFERENCE's own dsp/ headers (oracle/_ref/libsdrpp_ref.so, built by
`make -C oracle ref` from /root/reference where it lies). Run in the build container only; the vectors
are committed so the oracle port can be pinned on the GPU box, where /root/reference does not exist.
Inputs are regenerated from seeds (sdrpp_b200/synth.py); only outputs are stored."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from sdrpp_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


POST_CASES = [
    ("fm_nfm_lp", "fm", dict(sr=48e3, bw=12.5e3, low_pass=1, seed=11)),
    ("fm_wfm_nolp", "fm", dict(sr=250e3, bw=200e3, low_pass=0, seed=12)),
    ("am_off_gain", "am", dict(sr=24e3, bw=12e3, mode=0, attack=50.0 / 24e3, decay=5.0 / 24e3, dc=100.0 / 24e3, gain=3.0, seed=13)),
    ("am_carrier", "am", dict(sr=24e3, bw=12e3, mode=1, attack=50.0 / 24e3, decay=5.0 / 24e3, dc=100.0 / 24e3, gain=0.0, seed=14)),
    ("am_audio", "am", dict(sr=24e3, bw=12e3, mode=2, attack=50.0 / 24e3, decay=5.0 / 24e3, dc=100.0 / 24e3, gain=0.0, seed=15)),
    ("usb_agc", "ssb", dict(sr=48e3, bw=2.7e3, mode=0, agc=1, attack=50.0 / 48e3, decay=5.0 / 48e3, seed=16)),
    ("lsb_noagc", "ssb", dict(sr=48e3, bw=2.7e3, mode=1, agc=0, attack=50.0 / 48e3, decay=5.0 / 48e3, seed=17)),
]


def post_input(sr, seed):
    """VFO-rate test stream in ragged blocks: AM+FM modulated carrier with level steps (AGC attack, decay and the
    clip look-ahead all fire), plus noise."""
    rng = np.random.default_rng(seed)
    sizes = [int(sr / 200), 1, 777, int(sr / 200), int(sr / 100), 313]
    n = sum(sizes)
    t = np.arange(n) / sr
    m = 0.5 * np.sin(2 * np.pi * 700.0 * t) + 0.3 * np.sin(2 * np.pi * 1900.0 * t)
    level = np.where(t < 0.3 * t[-1], 0.3, np.where(t < 0.6 * t[-1], 6.0, 0.002))
    x = level * (1.0 + 0.8 * m) * np.exp(1j * (2 * np.pi * 300.0 * t + 2 * np.pi * 1500.0 * np.cumsum(m) / sr))
    x = x + 1e-3 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    x = x.astype(np.complex64)
    out, p = [], 0
    for s in sizes:
        out.append(x[p:p + s]); p += s
    return out


def make_post(lib, kind, a):
    if kind == "fm":
        return lib.fm_full(a["sr"], a["bw"], bool(a["low_pass"]))
    if kind == "am":
        return lib.am_full(a["mode"], a["bw"], a["attack"], a["decay"], a["dc"], a["sr"], a["gain"])
    return lib.ssb_full(a["mode"], a["bw"], a["sr"], bool(a["agc"]), a["attack"], a["decay"])


def wfm_input(n, sr, seed):
    """FM-modulated broadcast multiplex at the VFO rate: L / R tones, 19 kHz pilot, L-R on 38 kHz, a BPSK-like 57 kHz RDS
    subcarrier (1187.5 symbols/s), 75 kHz deviation."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    L, R = 0.5 * np.sin(2 * np.pi * 1000 * t), 0.3 * np.sin(2 * np.pi * 2500 * t)
    bits = rng.integers(0, 2, int(n / sr * 1187.5) + 2) * 2.0 - 1.0
    sym = bits[(t * 1187.5).astype(int)]
    mpx = (0.4 * (L + R) + 0.4 * (L - R) * np.sin(2 * np.pi * 38000 * t) + 0.1 * np.sin(2 * np.pi * 19000 * t)
           + 0.05 * sym * np.cos(2 * np.pi * 57000 * t))
    return (0.5 * np.exp(1j * 2 * np.pi * 75e3 * np.cumsum(mpx) / sr)).astype(np.complex64)


def pcm_input(n, seed):
    """Block whose negative excursion exceeds its (signed) maximum, so the compressor's saturation branch is taken."""
    rng = np.random.default_rng(seed)
    x = (0.3 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    x[n // 3] = np.complex64(-1.7 + 0.2j)
    return x


def if_input(nblocks, n, seed):
    """Blocks of a carrier that fades below and rises above a squelch level, with impulse noise for the blanker."""
    rng = np.random.default_rng(seed)
    t = np.arange(nblocks * n)
    env = np.where((t // n) % 16 < 3, 0.003, 0.4)          # three quiet blocks, thirteen loud ones, repeating
    x = env * np.exp(2j * np.pi * 0.013 * t) + 0.001 * (rng.standard_normal(len(t)) + 1j * rng.standard_normal(len(t)))
    x[rng.integers(0, len(t), len(t) // 97)] *= 25.0        # impulses
    return x.astype(np.complex64).reshape(nblocks, n)


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = po.Ref()
    cases = []

    def add(name, kind, args, arrays, **extra):
        f = name + ".npz"
        np.savez_compressed(os.path.join(GOLD, f), **arrays)
        cases.append(dict(name=name, kind=kind, args=list(args), file=f, **extra))

    rx = [("cfg1_wfm250", 2.4e6, 250e3, 200e3, 100e3, po.DEMOD_QUAD, [12000] * 3, 1),
          ("cfg1_wfm240", 2.4e6, 240e3, 200e3, -300e3, po.DEMOD_QUAD, [12000, 11999, 1, 12000], 1),
          ("cfg2_nfm", 3.2e6, 48e3, 12.5e3, 400e3, po.DEMOD_QUAD, [7936] * 5, 2),
          ("cfg3_wfm", 20e6, 250e3, 200e3, 3.1e6, po.DEMOD_QUAD, [100000] * 2, 3),
          ("cfg4_usb", 15.36e6, 48e3, 2.7e3, 1.0e6, po.DEMOD_USB, [76800] * 3, 4),
          ("cfg4_am", 15.36e6, 24e3, 12e3, -2.0e6, po.DEMOD_AM, [76800] * 3, 4),
          ("cfg5_nfm", 122.88e6, 48e3, 12.5e3, 30e6, po.DEMOD_QUAD, [614400], 5),
          ("cfg5_am", 122.88e6, 24e3, 12e3, -41e6, po.DEMOD_AM, [614400], 5),
          ("upsample", 48e3, 96e3, 96e3, 1e3, po.DEMOD_NONE, [480, 481], 6)]
    for name, inSR, outSR, bw, off, demod, blocks, seed in rx:
        n = sum(blocks)
        x = synth.baseband(n, inSR, seed, carriers=[(off, "fm")], noise_dbfs=-40.0).astype(np.complex64)
        v = ref.rxvfo(inSR, outSR, bw, off)
        d = ref.demod(demod, bw, outSR)
        ys, ds, p = [], [], 0
        for s in blocks:
            y = v.process(x[p:p + s]); p += s
            ys.append(y)
            if d is not None:
                ds.append(d.process(y))
        arrays = {"iq": np.concatenate(ys)}
        if d is not None:
            arrays["demod"] = np.concatenate(ds)
        add("rxvfo_" + name, "rxvfo", (inSR, outSR, bw, off, demod), arrays, n=n, seed=seed, blocks=blocks, counts=[len(y) for y in ys])

    for name, N, nz, wtype, seed in [("bh7_64k", 65536, 65536, po.WIN_BH7, 1), ("hann_8k_pad", 8192, 6000, po.WIN_HANN, 2),
                                     ("bh4_1k", 1024, 1024, po.WIN_BH4, 3)]:
        x = synth.baseband(nz, 2.4e6, seed, noise_dbfs=-40.0).astype(np.complex64)
        w = ref.window(wtype, nz)
        row32, _, row64 = ref.spectrum(N, x, w)
        add("spectrum_" + name, "spectrum", (N, nz, wtype), {"window": w, "row32": row32, "row64": row64}, seed=seed)

    # conversions: the reference does these inside its source modules, which cannot be compiled
    # here (vendor SDKs); the port's formulas are pinned by the SURVEY 8c hashes instead. The VOLK-based variants
    # go through the shim's (float)x/scale.
    x = synth.baseband(4 * 30720, 61.44e6, 7, noise_dbfs=-40.0).astype(np.complex64)
    pd, dc = ref.powerdecim(4), ref.dcblock(50.0 / (61.44e6 / 4))
    y = ref.conjugate(dc.process(pd.process(x)))
    add("frontend_x4_dc_conj", "frontend", (4,), {"out": y}, n=len(x), seed=7)

    # complete demodulators: dsp::demod::FM/AM/SSB<float> on a seeded VFO-rate stream
    for name, kind, args in POST_CASES:
        x = post_input(args["sr"], args["seed"])
        d = make_post(ref, kind, args)
        out = np.concatenate([d.process(b) for b in x])
        add("post_" + name, "post", (kind,), {"out": out}, params=args)

    # waterfall zoom (fft_scaler.h is self-contained, so the reference header itself generates these)
    for name, N, out, vo, vb, wb, seed in [("zoom_1m_max", 1048576, 1917, -2e7, 3.3e7, 122.88e6, 8), ("zoom_1k_point", 1024, 2000, 0.0, 2.4e6, 2.4e6, 9)]:
        rng = np.random.default_rng(seed)
        row = (rng.standard_normal(N) * 10.0 - 80.0).astype(np.float32)
        add("fft_" + name, "zoom", (N, out, vo, vb, wb), {"out": ref.fft_zoom(vo, vb, wb, row, out)}, seed=seed)

    # SDR++ server wire packets: the reference's own compressor makes the packet, its own
    # decompressor reads it back
    for name, ptype, n, seed in [("i8", 0, 4099, 10), ("i16", 1, 4099, 11), ("f32", 2, 513, 12)]:
        x = pcm_input(n, seed)
        pk = ref.pcm_compress(ptype, x)
        add("pcm_" + name, "pcm", (ptype, n), {"packet": pk, "out": ref.pcm_decompress(pk)}, seed=seed)

    # radio IF chain blocks: NoiseBlanker then Squelch on a bursty, fading stream, block by block
    x = if_input(20, 240, 13)
    nb, sq = ref.noise_blanker(500.0 / 48000.0, 3.0), ref.squelch(-20.0)
    y_nb = [nb.process(b) for b in x]
    y_sq = [sq.process(b) for b in y_nb]
    fm = ref.fm_if(15)
    y_fm = [fm.process(b) for b in y_sq]
    add("if_chain_nb_squelch", "if_chain", (500.0 / 48000.0, 3.0, -20.0, 20, 240),
        {"nb": np.concatenate(y_nb), "out": np.concatenate(y_sq), "fmif15": np.concatenate(y_fm)}, seed=13)

    # dsp::demod::BroadcastFM with its RDS side output: stereo audio and the 5 kS/s RDS baseband, ragged blocks
    for name, sr, stereo, seed in [("250k_stereo", 250e3, 1, 14), ("240k_mono", 240e3, 0, 15)]:
        blocks = [1250, 1, 777, 1250, 3000, 2, 1250, 1250, 1250, 1250]
        x = wfm_input(sum(blocks), sr, seed)
        d = ref.wfm(75e3, sr, stereo, 1, rds=True)
        lrs, rs, p = [], [], 0
        for b in blocks:
            lr, r = d.process(x[p:p + b]); p += b
            lrs.append(lr); rs.append(r)
        add("wfm_rds_" + name, "wfm_rds", (sr, stereo), {"lr": np.concatenate(lrs), "rds": np.concatenate(rs)}, seed=seed, blocks=blocks,
            counts=[len(r) for r in rs])

    json.dump({"generator": "tools/make_golden.py", "source": ref.lib.ref_build_info.restype and "oracle/_ref/libsdrpp_ref.so (reference dsp/ headers, IEEE flags)",
               "cases": cases}, open(os.path.join(GOLD, "manifest.json"), "w"), indent=1)
    print("wrote", len(cases), "cases;", sum(os.path.getsize(os.path.join(GOLD, c["file"])) for c in cases) // 1024, "KiB")


if __name__ == "__main__":
    main()
