"""GPU parity: post-detector stages of the three demodulators -- dsp::demod::FM<float> low-pass,
dsp::demod::AM<float> carrier/audio AGC + DC blocker + low-pass, dsp::demod::SSB<float> AGC -- through the C ABI.

The stages are checked in isolation: the oracle's complete demodulator is fed the GPU's own VFO output (cf32), so
the only differences left are those of the stages under test (the VFO itself is gated in test_gpu_channelizer.py).
Every branch is gated at 1e-5 relative RMS: the AGC and DC blocker recurrences and the discriminator's complex product use
the reference's operation order (atan2f implementations differ by ~1 ulp, 5e-8 here). The WFM stereo decoder
(dsp::demod::BroadcastFM) is checked the same way in all four stereo / low-pass modes."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

IN_SR, BLK = 2.4e6, 12000


def stream(nblocks, offset, seed):
    """A carrier at `offset` with AM + FM modulation and level steps (AGC attack, decay and clip look-ahead fire)."""
    rng = np.random.default_rng(seed)
    n = nblocks * BLK
    t = np.arange(n) / IN_SR
    m = 0.5 * np.sin(2 * np.pi * 700.0 * t) + 0.3 * np.sin(2 * np.pi * 1900.0 * t)
    level = np.where(t < 0.3 * t[-1], 0.02, np.where(t < 0.6 * t[-1], 0.5, 0.0005))
    ph = 2 * np.pi * ((offset * t) % 1.0) + 2 * np.pi * 1500.0 * np.cumsum(m) / IN_SR
    x = level * (1.0 + 0.8 * m) * np.exp(1j * ph) + 1e-5 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    x = x.astype(np.complex64)
    return [x[i * BLK:(i + 1) * BLK] for i in range(nblocks)]


def run(gpu, vfo, post, blocks, sizes=None):
    iq, audio = [], []
    with gpu.Frontend(IN_SR, max_block=BLK) as fe:
        vid = fe.add_vfo(*vfo)
        fe.set_post(vid, **post)
        for b in blocks:
            fe.process(po.FMT_CF32, b)
            y, _ = fe.vfo_output(vid)
            a = fe.vfo_audio(vid)
            assert len(a) == len(y)
            iq.append(y); audio.append(a)
    return iq, audio


CASES = [
    ("fm_nfm_lp", (48e3, 12.5e3, 100e3, po.DEMOD_QUAD), dict(fm_lowpass=True), lambda o: o.fm_full(48e3, 12.5e3, True), 1e-5),
    ("fm_wfm_lp", (250e3, 200e3, -300e3, po.DEMOD_QUAD), dict(fm_lowpass=True), lambda o: o.fm_full(250e3, 200e3, True), 1e-5),
    ("fm_nolp", (48e3, 12.5e3, 100e3, po.DEMOD_QUAD), dict(fm_lowpass=False), lambda o: o.fm_full(48e3, 12.5e3, False), 1e-5),
    ("am_off", (24e3, 12e3, 100e3, po.DEMOD_AM), dict(am_agc_mode=0, agc_attack=50 / 24e3, agc_decay=5 / 24e3, dc_block_rate=100 / 24e3, agc_gain=3.0),
     lambda o: o.am_full(0, 12e3, 50 / 24e3, 5 / 24e3, 100 / 24e3, 24e3, 3.0), 1e-5),
    ("am_carrier", (24e3, 12e3, 100e3, po.DEMOD_AM), dict(am_agc_mode=1, agc_attack=50 / 24e3, agc_decay=5 / 24e3, dc_block_rate=100 / 24e3),
     lambda o: o.am_full(1, 12e3, 50 / 24e3, 5 / 24e3, 100 / 24e3, 24e3), 1e-5),
    ("am_audio", (24e3, 12e3, 100e3, po.DEMOD_AM), dict(am_agc_mode=2, agc_attack=50 / 24e3, agc_decay=5 / 24e3, dc_block_rate=100 / 24e3),
     lambda o: o.am_full(2, 12e3, 50 / 24e3, 5 / 24e3, 100 / 24e3, 24e3), 1e-5),
    ("usb_agc", (48e3, 2.7e3, 100e3, po.DEMOD_USB), dict(ssb_agc=True, agc_attack=50 / 48e3, agc_decay=5 / 48e3),
     lambda o: o.ssb_full(0, 2.7e3, 48e3, True, 50 / 48e3, 5 / 48e3), 1e-5),
    ("lsb_fixed", (48e3, 2.7e3, 100e3, po.DEMOD_LSB), dict(ssb_agc=False, agc_attack=50 / 48e3, agc_decay=5 / 48e3, agc_gain=2.0),
     None, 1e-5),
]


@pytest.mark.parametrize("name,vfo,post,make,tol", CASES, ids=[c[0] for c in CASES])
def test_post_detector_stage(gpu, port, name, vfo, post, make, tol):
    blocks = stream(10, vfo[2], 21)
    iq, audio = run(gpu, vfo, post, blocks)
    if make is None:
        # SSB with the AGC off: fixed gain with clipping (agc.h:126-143) on the GPU's own front-end output
        with gpu.Frontend(IN_SR, max_block=BLK) as fe:
            vid = fe.add_vfo(*vfo)
            ref = []
            for b in blocks:
                fe.process(po.FMT_CF32, b)
                d = fe.vfo_output(vid)[1].astype(np.float32)
                g = np.float32(post["agc_gain"])
                amp = np.abs(d)
                with np.errstate(divide="ignore", invalid="ignore"):
                    ref.append(np.where(amp * g > np.float32(10.0), d * (np.float32(10.0) / amp), d * g).astype(np.float32))
    else:
        o = make(port)
        ref = [o.process(y) if len(y) else np.zeros(0, np.float32) for y in iq]
    ga, ra = np.concatenate(audio), np.concatenate(ref)
    assert len(ga) == len(ra) and len(ga) > 500
    assert np.all(np.isfinite(ga))
    if vfo[3] == po.DEMOD_QUAD:
        err = np.sqrt(np.mean((ga - ra) ** 2)) / max(np.sqrt(np.mean(ra ** 2)), 1e-3)
    else:
        err = po.rel_rms(ga, ra)
    assert err <= tol, f"{name}: rel-RMS {err:.3e}"


def test_post_follows_bandwidth_and_disable(gpu, port):
    """setBandwidth re-designs the demodulator's low-pass (am.h:46-55); disabling removes the audio output."""
    blocks = stream(6, 50e3, 22)
    with gpu.Frontend(IN_SR, max_block=BLK) as fe:
        vid = fe.add_vfo(24e3, 12e3, 50e3, po.DEMOD_AM)
        fe.set_post(vid, am_agc_mode=2, agc_attack=50 / 24e3, agc_decay=5 / 24e3, dc_block_rate=100 / 24e3)
        for b in blocks[:2]:
            fe.process(po.FMT_CF32, b)
        fe.vfo_set_bandwidth(vid, 8e3)
        o = port.am_full(2, 8e3, 50 / 24e3, 5 / 24e3, 100 / 24e3, 24e3)
        got, want = [], []
        for b in blocks[2:]:
            fe.process(po.FMT_CF32, b)
            y, _ = fe.vfo_output(vid)
            got.append(fe.vfo_audio(vid)); want.append(o.process(y))
        assert po.rel_rms(np.concatenate(got), np.concatenate(want)) <= 1e-5
        fe.set_post(vid, enabled=False)
        fe.process(po.FMT_CF32, blocks[0])
        with pytest.raises(gpu.SdrppCudaError):
            fe.vfo_audio(vid)


def test_post_many_vfos_mixed(gpu, port):
    """Post stages on a subset of a mixed VFO set: only those VFOs get audio, the others are untouched."""
    blocks = stream(4, 100e3, 23)
    vf = [(48e3, 12.5e3, 100e3, po.DEMOD_QUAD), (24e3, 12e3, 100e3, po.DEMOD_AM), (48e3, 2.7e3, 100e3, po.DEMOD_USB),
          (48e3, 12.5e3, -200e3, po.DEMOD_QUAD), (24e3, 12e3, 100e3, po.DEMOD_AM)]
    with gpu.Frontend(IN_SR, max_block=BLK) as fe:
        ids = [fe.add_vfo(*v) for v in vf]
        fe.set_post(ids[0], fm_lowpass=True)
        fe.set_post(ids[1], am_agc_mode=1, agc_attack=50 / 24e3, agc_decay=5 / 24e3, dc_block_rate=100 / 24e3)
        fe.set_post(ids[2], ssb_agc=True, agc_attack=50 / 48e3, agc_decay=5 / 48e3)
        os_ = [port.fm_full(48e3, 12.5e3, True), port.am_full(1, 12e3, 50 / 24e3, 5 / 24e3, 100 / 24e3, 24e3),
               port.ssb_full(0, 2.7e3, 48e3, True, 50 / 48e3, 5 / 48e3)]
        got, want = [[], [], []], [[], [], []]
        for b in blocks:
            fe.process(po.FMT_CF32, b)
            for k in range(3):
                y, _ = fe.vfo_output(ids[k])
                got[k].append(fe.vfo_audio(ids[k])); want[k].append(os_[k].process(y))
            with pytest.raises(gpu.SdrppCudaError):
                fe.vfo_audio(ids[3])
        for k, tol in ((0, 1e-5), (1, 1e-5), (2, 1e-5)):
            g, w = np.concatenate(got[k]), np.concatenate(want[k])
            err = np.sqrt(np.mean((g - w) ** 2)) / max(np.sqrt(np.mean(w ** 2)), 1e-3)
            assert err <= tol, (k, err)


def wfm_stream(nblocks, offset, seed):
    """A stereo FM broadcast signal: L = 1 kHz, R = 2.5 kHz, 19 kHz pilot, L-R on the 38 kHz subcarrier, 75 kHz deviation."""
    rng = np.random.default_rng(seed)
    n = nblocks * BLK
    t = np.arange(n) / IN_SR
    L, R = 0.5 * np.sin(2 * np.pi * 1000.0 * t), 0.3 * np.sin(2 * np.pi * 2500.0 * t)
    mpx = 0.45 * (L + R) + 0.45 * (L - R) * np.sin(2 * np.pi * 38000.0 * t) + 0.1 * np.sin(2 * np.pi * 19000.0 * t)
    ph = 2 * np.pi * ((offset * t) % 1.0) + 2 * np.pi * 75e3 * np.cumsum(mpx) / IN_SR
    x = 0.5 * np.exp(1j * ph) + 1e-4 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    x = x.astype(np.complex64)
    return [x[i * BLK:(i + 1) * BLK] for i in range(nblocks)]


@pytest.mark.parametrize("out_sr", [250e3, 240e3])
@pytest.mark.parametrize("stereo,low_pass", [(True, True), (True, False), (False, True), (False, False)])
def test_wfm_stereo_decoder(gpu, port, report, out_sr, stereo, low_pass):
    """dsp::demod::BroadcastFM behind a WFM VFO (radio module: bandwidth 150 kHz, deviation = bandwidth / 2): the oracle's
    decoder applied to the GPU's own VFO output against the GPU's (l, r), <= 1e-5 -- the pilot PLL is a feedback loop over
    cosf / sinf / atan2f, whose implementations differ by an ulp, and stays inside the gate -- and the two channels separate."""
    bw = 150e3
    blocks = wfm_stream(40, -300e3, 27)
    o = port.wfm(bw / 2.0, out_sr, stereo, low_pass)
    got, want = [], []
    with gpu.Frontend(IN_SR, max_block=BLK) as fe:
        vid = fe.add_vfo(out_sr, bw, -300e3, po.DEMOD_QUAD)
        fe.set_post(vid, fm_lowpass=low_pass, wfm=True, wfm_stereo=stereo)
        for b in blocks:
            fe.process(po.FMT_CF32, b)
            y, _ = fe.vfo_output(vid)
            l, r = fe.vfo_audio_stereo(vid)
            assert len(l) == len(r) == len(y)
            got.append(np.stack([l, r], axis=1)); want.append(o.process(y))
    g, w = np.concatenate(got), np.concatenate(want)
    assert g.shape == w.shape and np.all(np.isfinite(g))
    s = len(g) // 4                               # past the PLL's lock-in and the filters' start-up
    err = float(np.sqrt(np.mean((g[s:].astype(np.float64) - w[s:]) ** 2)) / np.sqrt(np.mean(w[s:].astype(np.float64) ** 2)))
    err0 = float(np.sqrt(np.mean((g[:s].astype(np.float64) - w[:s]) ** 2)) / np.sqrt(np.mean(w[:s].astype(np.float64) ** 2)))
    report(f"8f-4 BroadcastFM {out_sr/1e3:g}k stereo={int(stereo)} lp={int(low_pass)}", locked=err, lock_in_quarter=err0, gate=1e-5)
    assert err <= 1e-5, f"rel-RMS {err:.3e}"
    assert err0 <= 1e-4, f"lock-in: {err0:.3e}"
    if stereo and low_pass:
        # channel separation: 1 kHz on the left only, 2.5 kHz on the right only
        n = len(g) - s
        sp = np.abs(np.fft.rfft(g[s:] * np.hanning(n)[:, None], axis=0))
        f = np.fft.rfftfreq(n, 1.0 / out_sr)
        k1, k2 = int(np.argmin(np.abs(f - 1000.0))), int(np.argmin(np.abs(f - 2500.0)))
        assert sp[k1, 0] > 10.0 * sp[k1, 1] and sp[k2, 1] > 10.0 * sp[k2, 0]   # >= 20 dB (the reference decoder itself gives 26 dB at 240 kS/s)
    if not stereo:
        assert np.array_equal(g[:, 0], g[:, 1])


def wfm_rds_stream(nblocks, offset, seed):
    """wfm_stream plus an RDS-like BPSK subcarrier at 57 kHz (1187.5 symbols/s)."""
    rng = np.random.default_rng(seed)
    n = nblocks * BLK
    t = np.arange(n) / IN_SR
    L, R = 0.5 * np.sin(2 * np.pi * 1000.0 * t), 0.3 * np.sin(2 * np.pi * 2500.0 * t)
    bits = rng.integers(0, 2, int(t[-1] * 1187.5) + 2) * 2.0 - 1.0
    sym = bits[(t * 1187.5).astype(int)]
    mpx = (0.4 * (L + R) + 0.4 * (L - R) * np.sin(2 * np.pi * 38000.0 * t) + 0.1 * np.sin(2 * np.pi * 19000.0 * t)
           + 0.05 * sym * np.cos(2 * np.pi * 57000.0 * t))
    ph = 2 * np.pi * ((offset * t) % 1.0) + 2 * np.pi * 75e3 * np.cumsum(mpx) / IN_SR
    x = 0.5 * np.exp(1j * ph) + 1e-4 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    x = x.astype(np.complex64)
    return [x[i * BLK:(i + 1) * BLK] for i in range(nblocks)]


@pytest.mark.parametrize("out_sr", [250e3, 240e3])
@pytest.mark.parametrize("stereo", [True, False])
def test_wfm_rds_side_output(gpu, port, report, out_sr, stereo):
    """BroadcastFM::rdsOut (broadcast_fm.h:168-175,188-198): (mpx, 0) -> FrequencyXlator(-57 kHz) -> RationalResampler(-> 5 kS/s).
    Stage-isolated: the oracle's translation (ideal-NCO flavour) and resampler applied to the GPU's own
    discriminator row; per-block sample counts exact, samples <= 1e-5. The reference decoder's own RDS output (fp32 rotator)
    on the GPU's VFO output is reported beside it. Ragged blocks; two VFOs share the plan (250 k: a 5000-phase bank)."""
    bw = 150e3
    blocks = wfm_rds_stream(60, -300e3, 31)
    sizes = [BLK, 7, BLK - 1, 5000, BLK, 1, 11999]
    xl = [port.xlator(-57000.0, out_sr, ideal=True) for _ in range(2)]
    rr = [port.resampler(out_sr, 5000.0) for _ in range(2)]
    full = port.wfm(bw / 2.0, out_sr, stereo, True, rds=True)
    got, want, ref_own = [[], []], [[], []], []
    audio_g, audio_w = [], []
    with gpu.Frontend(IN_SR, max_block=BLK) as fe:
        ids = [fe.add_vfo(out_sr, bw, -300e3, po.DEMOD_QUAD), fe.add_vfo(out_sr, bw, -300e3 + 20e3, po.DEMOD_QUAD)]
        for vid in ids:
            fe.set_post(vid, fm_lowpass=True, wfm=True, wfm_stereo=stereo, wfm_rds=True)
        other = fe.add_vfo(out_sr, bw, 100e3, po.DEMOD_QUAD)
        fe.set_post(other, fm_lowpass=True, wfm=True, wfm_stereo=stereo)
        for i, b in enumerate(blocks):
            b = b[:sizes[i % len(sizes)]]
            fe.process(po.FMT_CF32, b)
            for k, vid in enumerate(ids):
                y, dm = fe.vfo_output(vid)
                r = fe.vfo_rds(vid)
                w = rr[k].process(xl[k].process(dm.astype(np.complex64)))
                assert len(r) == len(w), (i, k, len(r), len(w))
                got[k].append(r); want[k].append(w)
                if k == 0:
                    lr, r_own = full.process(y)
                    assert len(r_own) == len(r)
                    ref_own.append(r_own)
                    l, rt = fe.vfo_audio_stereo(vid)
                    audio_g.append(np.stack([l, rt], axis=1)); audio_w.append(lr)
            with pytest.raises(gpu.SdrppCudaError):
                fe.vfo_rds(other)
    for k in range(2):
        g, w = np.concatenate(got[k]), np.concatenate(want[k])
        assert len(g) > 500 and np.all(np.isfinite(g.view(np.float32)))
        err = float(np.sqrt(np.sum(np.abs(g.astype(np.complex128) - w) ** 2) / np.sum(np.abs(w.astype(np.complex128)) ** 2)))
        if k == 0:
            o = np.concatenate(ref_own)
            walk = float(np.sqrt(np.sum(np.abs(o.astype(np.complex128) - g) ** 2) / np.sum(np.abs(g.astype(np.complex128)) ** 2)))
            report(f"8f-4 BroadcastFM RDS output {out_sr/1e3:g}k stereo={int(stereo)}", stage_isolated_vs_ideal_nco=err,
                   reference_decoder_fp32_rotator_vs_gpu=walk, samples=len(g), gate=1e-5)
            assert walk <= 2e-4, walk
        assert err <= 1e-5, f"vfo {k}: rel-RMS {err:.3e}"
    # the audio outputs are unchanged by the side output
    ag, aw = np.concatenate(audio_g), np.concatenate(audio_w)
    s = len(ag) // 4
    aerr = float(np.sqrt(np.mean((ag[s:].astype(np.float64) - aw[s:]) ** 2)) / np.sqrt(np.mean(aw[s:].astype(np.float64) ** 2)))
    assert aerr <= 1e-5, aerr
