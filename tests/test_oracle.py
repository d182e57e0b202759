"""Pins the oracle. The reference has no tests or golden vectors for this path, so
the plain-C port (oracle/port/oracle.c) is pinned (a) bit-exactly against the reference's own dsp/
headers compiled in oracle/_ref (when present: this container) and (b) against golden vectors
generated from that build and committed under tests/golden/ (tools/make_golden.py), which is what
travels to the GPU box. Known answers from SURVEY 8c are checked too."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from sdrpp_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def _noise(n, seed):
    rng = np.random.default_rng(seed)
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64) * np.float32(0.25)


# ---- known answers ---------------------------------------------------------------------
def test_conversion_known_answers(port):
    t = port.convert(po.FMT_U8_RTL, np.repeat(np.arange(256, dtype=np.uint8), 2)).view(np.float32)[0::2]
    assert t[0] == np.float32(-1.0) and t[255] == np.float32(1.0) and t[128] == np.float32(0.00392156886)
    assert po.fnv1a_words(t) == 0xb8121cc5
    t16 = port.convert(po.FMT_I16_FILE, np.repeat(np.arange(-32768, 32768, dtype=np.int32).astype(np.int16), 2)).view(np.float32)[0::2]
    assert po.fnv1a_words(t16) == 0x4fdefa45


def test_unity_gain_tone(port):
    N, k = 65536, 321
    x = np.exp(2j * np.pi * k * np.arange(N) / N).astype(np.complex64)
    row32, _, row64 = port.spectrum(N, x, port.window(po.WIN_BH7, N))
    assert int(np.argmax(row64)) == N // 2 + k and abs(row64[N // 2 + k]) < 1e-3 and abs(row32[N // 2 + k]) < 2e-3


def test_rxvfo_known_answer(port):
    n = np.arange(12000)
    x = np.exp(2j * np.pi * 100e3 * n / 2.4e6).astype(np.complex64)
    v = port.rxvfo(2.4e6, 240e3, 200e3, 100e3)
    y = v.process(x)
    assert len(y) == 1200
    assert abs(y[-1].real - 0.996594) < 2e-4 and abs(y[-1].imag) < 1e-4
    q = port.quadrature(100e3, 240e3).process(y)
    assert abs(q[-1]) < 1e-4


def test_block_counts(port):
    v = port.rxvfo(3.2e6, 48e3, 12.5e3, 0.0)
    counts = [len(v.process(_noise(7936, i))) for i in range(40)]
    assert counts[:3] == [120, 119, 119] and sum(counts) == 4762


# ---- port vs the compiled reference (this container only) -----------------------------------------------
@pytest.mark.parametrize("cfg", [(2.4e6, 250e3, 200e3, 100e3, 12000), (3.2e6, 48e3, 12.5e3, -400e3, 7936),
                                 (20e6, 240e3, 200e3, 3.1e6, 100000), (15.36e6, 48e3, 2.7e3, 1e6, 76800), (48e3, 96e3, 96e3, 1e3, 480)])
def test_port_matches_reference_rxvfo(port, ref, cfg):
    inSR, outSR, bw, off, blk = cfg
    a, b = port.rxvfo(inSR, outSR, bw, off), ref.rxvfo(inSR, outSR, bw, off)
    assert a.info() == b.info()
    for i in range(3):
        x = _noise(blk if i != 1 else blk - 37, 100 + i)
        ya, yb = a.process(x), b.process(x)
        assert len(ya) == len(yb) and np.array_equal(_bits(ya), _bits(yb))


def test_port_matches_reference_pieces(port, ref):
    x = _noise(30000, 1)
    for ratio in (2, 4, 64, 2048):
        a, b = port.powerdecim(ratio), ref.powerdecim(ratio)
        for i in range(3):
            xs = x[i * 9001:(i + 1) * 9001]
            assert np.array_equal(_bits(a.process(xs)), _bits(b.process(xs)))
            assert a.offsets() == b.offsets()
    taps = port.lowpass_taps(12e3, 1.2e3, 5 * 48e3) * np.float32(5)
    a, b = port.polyphase(5, 6, taps), ref.polyphase(5, 6, taps)
    for i in range(3):
        xs = x[i * 777:(i + 1) * 777 + i]
        assert np.array_equal(_bits(a.process(xs)), _bits(b.process(xs))) and a.state() == b.state()
    a, b = port.dcblock(50.0 / 2.4e6), ref.dcblock(50.0 / 2.4e6)
    assert np.array_equal(_bits(a.process(x + 0.1)), _bits(b.process(x + 0.1)))
    for dem in (po.DEMOD_QUAD, po.DEMOD_AM, po.DEMOD_USB, po.DEMOD_LSB):
        a, b = port.demod(dem, 12.5e3, 48e3), ref.demod(dem, 12.5e3, 48e3)
        assert np.array_equal(_bits(a.process(x[:4000])), _bits(b.process(x[:4000])))
    w = port.window(po.WIN_BH4, 4096)
    ra, rb = port.spectrum(4096, x, w), ref.spectrum(4096, x, w)
    assert np.array_equal(_bits(ra[0]), _bits(rb[0])) and np.array_equal(ra[1], rb[1])


def test_port_matches_reference_full_demods(port, ref):
    """dsp::demod::FM/AM/SSB<float> including AGC (clip look-ahead), DC blocker and the audio low-pass."""
    from tools.make_golden import POST_CASES, make_post, post_input
    for name, kind, a in POST_CASES:
        da, db = make_post(port, kind, a), make_post(ref, kind, a)
        for blk in post_input(a["sr"], a["seed"] + 100):
            assert np.array_equal(_bits(da.process(blk)), _bits(db.process(blk))), name


# ---- SDR++ server wire packets (dsp/compression): port vs the reference headers ------------------------------
@pytest.mark.parametrize("ptype", [0, 1, 2])
@pytest.mark.parametrize("n", [1, 2, 5, 1000, 30001])
def test_pcm_port_vs_ref(port, ref, ptype, n):
    from tools.make_golden import pcm_input
    x = pcm_input(n, 1000 * ptype + n)
    pr, pp = ref.pcm_compress(ptype, x), port.pcm_compress(ptype, x)
    assert len(pr) == 8 + n * (2, 4, 8)[ptype]
    assert np.array_equal(pr, pp)
    assert np.array_equal(_bits(ref.pcm_decompress(pr)), _bits(port.pcm_decompress(pr)))
    # header: compression type 0, sample type, scaler = largest signed scalar (0 for float32)
    assert pr[:4].view(np.uint16).tolist() == [0, ptype]
    assert pr[4:8].view(np.float32)[0] == (np.float32(0) if ptype == 2 else x.view(np.float32).max())


def test_pcm_unknown_type_and_ragged_payload(port, ref):
    pk = np.zeros(8 + 7, dtype=np.uint8)
    pk[2] = 7  # unknown PCMType: the decompressor emits nothing
    assert len(port.pcm_decompress(pk)) == 0 and len(ref.pcm_decompress(pk)) == 0
    pk[2] = 1  # int16 with a payload that is not a whole number of samples: floor((count-8)/4)
    pk[4:8] = np.array([0.5], np.float32).view(np.uint8)
    pk[8:] = np.arange(7, dtype=np.uint8)
    a, b = ref.pcm_decompress(pk), port.pcm_decompress(pk)
    assert len(a) == 1 and np.array_equal(_bits(a), _bits(b))


# ---- radio IF chain blocks: port vs the reference headers ----------------------------------------------------
@pytest.mark.parametrize("n", [1, 7, 240, 1250])
def test_if_chain_port_vs_ref(port, ref, n):
    from tools.make_golden import if_input
    x = if_input(40, n, 100 + n)
    for level in (2.0, 10.0):
        a, b = ref.noise_blanker(500.0 / 24000.0, level), port.noise_blanker(500.0 / 24000.0, level)
        for blk in x:
            assert np.array_equal(_bits(a.process(blk)), _bits(b.process(blk)))
    for sq_level in (-60.0, -20.0, -8.5, 0.0):
        a, b = ref.squelch(sq_level), port.squelch(sq_level)  # one reference Squelch at a time: its counter is a static
        for blk in x:
            assert np.array_equal(_bits(a.process(blk)), _bits(b.process(blk)))
        del a


@pytest.mark.parametrize("bins", [9, 15, 31, 32])
def test_fm_if_port_vs_ref(port, ref, bins):
    """FMIF (noise_reduction/fm_if.h) over the DFT stand-in for FFTW: port and reference headers agree bit for bit."""
    from tools.make_golden import if_input
    a, b = ref.fm_if(bins), port.fm_if(bins)
    for blk in if_input(5, 150, 70 + bins):
        assert np.array_equal(_bits(a.process(blk)), _bits(b.process(blk)))


def test_fm_if_tone_passes_with_window_gain(port):
    """Known answer: a tone on bin 3 of 32 comes out as the same tone scaled by the window's sum, delayed by bins/2 - 1
    samples -- the backward DFT's element bins/2 re-centres the phase at the middle of the window."""
    bins, n = 32, 256
    t = np.arange(n + bins)
    x = np.exp(2j * np.pi * 3 * t / bins).astype(np.complex64)
    f = port.fm_if(bins)
    y = f.process(x)[bins:]          # past the start-up (zero history)
    k = np.arange(bins)
    w = sum(c * s * np.cos(2 * np.pi * i * k / (bins - 1)) for i, (c, s) in enumerate(zip([0.355768, 0.487396, 0.144232, 0.012604], [1, -1, 1, -1])))
    want = w.sum() * x[bins - (bins // 2 - 1):][:len(y)]
    assert np.max(np.abs(y - want)) < 2e-5 * w.sum()


# ---- golden vectors generated from the compiled reference (travel to the GPU box) -----------------------------
def _golden_cases():
    p = os.path.join(GOLD, "manifest.json")
    return json.load(open(p))["cases"] if os.path.exists(p) else []


@pytest.mark.parametrize("case", _golden_cases(), ids=lambda c: c["name"])
def test_port_matches_golden(port, case):
    data = np.load(os.path.join(GOLD, case["file"]))
    kind = case["kind"]
    if kind == "rxvfo":
        inSR, outSR, bw, off, demod = case["args"]
        x = synth.baseband(case["n"], inSR, case["seed"], carriers=[(off, "fm")], noise_dbfs=-40.0).astype(np.complex64)
        v = port.rxvfo(inSR, outSR, bw, off)
        d = port.demod(demod, bw, outSR)
        ys, ds, p = [], [], 0
        for s in case["blocks"]:
            y = v.process(x[p:p + s]); p += s
            ys.append(y)
            if d is not None:
                ds.append(d.process(y))
        assert [len(y) for y in ys] == case["counts"]
        assert np.array_equal(_bits(np.concatenate(ys)), _bits(data["iq"]))
        if d is not None:
            assert np.array_equal(_bits(np.concatenate(ds)), _bits(data["demod"]))
    elif kind == "spectrum":
        N, nz, wtype = case["args"]
        x = synth.baseband(nz, 2.4e6, case["seed"], noise_dbfs=-40.0).astype(np.complex64)
        w = port.window(wtype, nz)
        assert np.array_equal(_bits(w), _bits(data["window"]))
        row32, _, row64 = port.spectrum(N, x, w)
        assert np.array_equal(_bits(row32), _bits(data["row32"]))
        assert np.allclose(row64, data["row64"], rtol=0, atol=1e-9)
    elif kind == "convert":
        raw = data["raw"]
        assert np.array_equal(_bits(port.convert(case["args"][0], raw)), _bits(data["out"]))
    elif kind == "frontend":
        ratio, = case["args"]
        x = synth.baseband(case["n"], 61.44e6, case["seed"], noise_dbfs=-40.0).astype(np.complex64)
        pd, dc = port.powerdecim(ratio), port.dcblock(50.0 / (61.44e6 / ratio))
        y = port.conjugate(dc.process(pd.process(x)))
        assert np.array_equal(_bits(y), _bits(data["out"]))
    elif kind == "zoom":
        N, out, vo, vb, wb = case["args"]
        rng = np.random.default_rng(case["seed"])
        row = (rng.standard_normal(int(N)) * 10.0 - 80.0).astype(np.float32)
        got, _ = port.fft_zoom(vo, vb, wb, row, int(out))
        assert np.array_equal(_bits(got), _bits(data["out"]))
    elif kind == "post":
        from tools.make_golden import make_post, post_input
        a = case["params"]
        d = make_post(port, case["args"][0], a)
        got = np.concatenate([d.process(b) for b in post_input(a["sr"], a["seed"])])
        assert np.array_equal(_bits(got), _bits(data["out"]))
    elif kind == "pcm":
        from tools.make_golden import pcm_input
        ptype, n = case["args"]
        pk = port.pcm_compress(ptype, pcm_input(n, case["seed"]))
        assert np.array_equal(pk, data["packet"])
        assert np.array_equal(_bits(port.pcm_decompress(data["packet"])), _bits(data["out"]))
    elif kind == "if_chain":
        from tools.make_golden import if_input
        rate, lvl, sq_level, nblocks, n = case["args"]
        nb, sq = port.noise_blanker(rate, lvl), port.squelch(sq_level)
        y_nb = [nb.process(b) for b in if_input(int(nblocks), int(n), case["seed"])]
        y_sq = [sq.process(b) for b in y_nb]
        assert np.array_equal(_bits(np.concatenate(y_nb)), _bits(data["nb"]))
        assert np.array_equal(_bits(np.concatenate(y_sq)), _bits(data["out"]))
        fm = port.fm_if(15)
        assert np.array_equal(_bits(np.concatenate([fm.process(b) for b in y_sq])), _bits(data["fmif15"]))
        muted = [not b.any() for b in y_sq]
        assert any(muted) and not all(muted)  # the fixture exercises both squelch states
    elif kind == "wfm_rds":
        from tools.make_golden import wfm_input
        sr, stereo = case["args"]
        x = wfm_input(sum(case["blocks"]), sr, case["seed"])
        d = port.wfm(75e3, sr, int(stereo), 1, rds=True)
        lrs, rs, p = [], [], 0
        for b in case["blocks"]:
            lr, r = d.process(x[p:p + b]); p += b
            lrs.append(lr); rs.append(r)
        assert [len(r) for r in rs] == case["counts"]
        assert np.array_equal(_bits(np.concatenate(lrs)), _bits(data["lr"]))
        assert np.array_equal(_bits(np.concatenate(rs)), _bits(data["rds"]))
    else:
        pytest.fail("unknown golden kind " + kind)


def test_golden_present():
    assert len(_golden_cases()) >= 8, "tests/golden is empty: run tools/make_golden.py where /root/reference exists"


@pytest.mark.parametrize("stereo,low_pass", [(1, 1), (1, 0), (0, 1), (0, 0)])
def test_port_broadcast_fm_bit_exact_vs_reference(port, ref, stereo, low_pass):
    """dsp::demod::BroadcastFM (demod/broadcast_fm.h): the C restatement against the reference's own header, block by block."""
    sr, n, blk = 250e3, 25000, 1250
    t = np.arange(n) / sr
    L, R = 0.5 * np.sin(2 * np.pi * 1000 * t), 0.3 * np.sin(2 * np.pi * 2500 * t)
    mpx = 0.45 * (L + R) + 0.45 * (L - R) * np.sin(2 * np.pi * 38000 * t) + 0.1 * np.sin(2 * np.pi * 19000 * t)
    x = (0.5 * np.exp(1j * 2 * np.pi * 75e3 * np.cumsum(mpx) / sr)).astype(np.complex64)
    a, b = port.wfm(75e3, sr, stereo, low_pass), ref.wfm(75e3, sr, stereo, low_pass)
    for i in range(0, n, blk):
        ya, yb = a.process(x[i:i + blk]), b.process(x[i:i + blk])
        assert ya.shape == (blk, 2) and np.array_equal(ya.view(np.uint32), yb.view(np.uint32))


def _wfm_rds_signal(sr, n, seed=5):
    from tools.make_golden import wfm_input
    return wfm_input(n, sr, seed)


@pytest.mark.parametrize("sr", [250e3, 240e3])
@pytest.mark.parametrize("stereo", [1, 0])
def test_port_broadcast_fm_rds_bit_exact_vs_reference(port, ref, sr, stereo):
    """BroadcastFM with _rdsOut (broadcast_fm.h:168-175,188-198): the decoder's own xlator(-57 kHz) + RationalResampler(-> 5 kS/s)
    on (mpx, 0); ragged blocks, audio and RDS outputs and the per-call RDS counts bit for bit."""
    n = 30000
    x = _wfm_rds_signal(sr, n)
    a, b = port.wfm(75e3, sr, stereo, 1, rds=True), ref.wfm(75e3, sr, stereo, 1, rds=True)
    sizes = [1250, 1, 777, 1250, 3000, 2, 1250]
    i, k, total = 0, 0, 0
    while i < n:
        m = min(sizes[k % len(sizes)], n - i); k += 1
        (ya, ra), (yb, rb) = a.process(x[i:i + m]), b.process(x[i:i + m])
        assert np.array_equal(ya.view(np.uint32), yb.view(np.uint32))
        assert len(ra) == len(rb) and np.array_equal(ra.view(np.uint32), rb.view(np.uint32))
        total += len(ra)
        i += m
    assert abs(total - n * 5000.0 / sr) <= 2 + n * 1e-4     # 250 k: IntSR rounds to 7813, interp/decim = 5000/7813


def test_port_rds_ideal_flavour_close_to_rotator(port):
    """The ideal-NCO flavour of the RDS translation against the reference rotator on a short run: the rotator has barely
    walked yet, so the two agree to ~1e-6; everything behind the translation is the same code."""
    sr, n = 250e3, 25000
    x = _wfm_rds_signal(sr, n)
    a, b = port.wfm(75e3, sr, 1, 1, rds=True), port.wfm(75e3, sr, 1, 1, rds=True, ideal_nco=True)
    (_, ra), (_, rb) = a.process(x), b.process(x)
    assert len(ra) == len(rb) > 400
    err = np.sqrt(np.sum(np.abs(ra - rb) ** 2) / np.sum(np.abs(rb) ** 2))
    assert err < 2e-5, err
