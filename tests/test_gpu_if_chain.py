"""GPU parity: the radio IF chain between the VFO output and the demodulator front end --
dsp::noise_reduction::NoiseBlanker -> dsp::noise_reduction::Squelch -> dsp::noise_reduction::FMIF (decoder_modules/radio/src/radio_module.h:73-78)
-- through the C ABI.

The blocks are checked in isolation, like the post-detector stages: the oracle's NoiseBlanker, Squelch and demod
front end are fed the GPU's own VFO output, so what is compared is the chain itself. Both blocks are recurrences /
in-order sums executed with the reference's operation order, so the AM (magnitude) gate is 1e-6 relative RMS and the
quadrature gate the discriminator's 2e-4 (atan2f implementations differ by ~1 ulp)."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

IN_SR, BLK = 2.4e6, 12000


def stream(nblocks, offset, seed):
    """Carrier at `offset` that fades under and rises over the squelch level every few blocks, FM + AM modulated,
    with short strong bursts (the blanker's excess test fires) on a weak noise floor."""
    rng = np.random.default_rng(seed)
    n = nblocks * BLK
    t = np.arange(n) / IN_SR
    blk = np.arange(n) // BLK
    level = np.where(blk % 24 < 5, 0.0008, 0.3)
    m = 0.5 * np.sin(2 * np.pi * 700.0 * t)
    ph = 2 * np.pi * ((offset * t) % 1.0) + 2 * np.pi * 2000.0 * np.cumsum(m) / IN_SR
    x = level * (1.0 + 0.5 * m) * np.exp(1j * ph) + 1e-4 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    for s in rng.integers(0, n - 400, nblocks // 2):   # bursts of ~150 us on the carrier frequency
        x[s:s + 360] += 3.0 * np.exp(1j * ph[s:s + 360])
    x = x.astype(np.complex64)
    return [x[i * BLK:(i + 1) * BLK] for i in range(nblocks)]


def rel_rms(a, b):
    a, b = np.asarray(a), np.asarray(b)
    a, b = (a.astype(np.complex128), b.astype(np.complex128)) if np.iscomplexobj(a) or np.iscomplexobj(b) else (a.astype(np.float64), b.astype(np.float64))
    return float(np.sqrt(np.sum(np.abs(a - b) ** 2) / max(np.sum(np.abs(b) ** 2), 1e-300)))


def front_end(port, demod, bw, sr):
    if demod == po.DEMOD_AM:
        return lambda y: port.am_magnitude(y)
    q = port.quadrature(bw / 2.0, sr)
    return lambda y: q.process(y)


CASES = [
    ("am_nb", (24e3, 12e3, 100e3, po.DEMOD_AM), dict(nb=True, nb_rate=500.0 / 24e3, nb_level=1.6), 1e-6),
    ("am_squelch", (24e3, 12e3, 100e3, po.DEMOD_AM), dict(squelch=True, squelch_level=-30.0), 1e-6),
    ("am_nb_squelch", (24e3, 12e3, -300e3, po.DEMOD_AM), dict(nb=True, nb_rate=500.0 / 24e3, nb_level=1.6, squelch=True, squelch_level=-30.0), 1e-6),
    ("nfm_nb_squelch", (48e3, 12.5e3, 100e3, po.DEMOD_QUAD), dict(nb=True, nb_rate=500.0 / 48e3, nb_level=1.6, squelch=True, squelch_level=-30.0), 2e-4),
    ("nfm_fmif15", (48e3, 12.5e3, 100e3, po.DEMOD_QUAD), dict(fmif_bins=15), 2e-4),
    ("nfm_all_fmif31", (48e3, 12.5e3, -300e3, po.DEMOD_QUAD), dict(nb=True, nb_rate=500.0 / 48e3, nb_level=1.6, squelch=True, squelch_level=-30.0, fmif_bins=31), 2e-4),
    ("am_fmif9", (24e3, 12e3, 100e3, po.DEMOD_AM), dict(fmif_bins=9), 1e-5),
    ("wfm_fmif32", (250e3, 200e3, -300e3, po.DEMOD_QUAD), dict(fmif_bins=32), 2e-4),
    ("wfm_squelch", (250e3, 200e3, -300e3, po.DEMOD_QUAD), dict(squelch=True, squelch_level=-30.0), 2e-4),
]


@pytest.mark.parametrize("name,vfo,chain,tol", CASES, ids=[c[0] for c in CASES])
def test_if_chain(gpu, port, name, vfo, chain, tol):
    blocks = stream(60, vfo[2], 31)
    out_sr, bw, _, demod = vfo
    nb = port.noise_blanker(chain["nb_rate"], chain["nb_level"]) if chain.get("nb") else None
    sq = port.squelch(chain["squelch_level"]) if chain.get("squelch") else None
    fm = port.fm_if(chain["fmif_bins"]) if chain.get("fmif_bins") else None
    fe_ref = front_end(port, demod, bw, out_sr)
    got, want, raw, muted_gpu, muted_ref = [], [], [], [], []
    with gpu.Frontend(IN_SR, max_block=BLK) as fe, gpu.Frontend(IN_SR, max_block=BLK) as fe_raw:
        vid = fe.add_vfo(*vfo)
        vraw = fe_raw.add_vfo(*vfo)
        fe.set_if_chain(vid, **chain)
        for b in blocks:
            fe.process(po.FMT_CF32, b)
            fe_raw.process(po.FMT_CF32, b)
            y, d = fe.vfo_output(vid)
            y0, d0 = fe_raw.vfo_output(vraw)
            # the iq result stays the raw VFO output (the plain VFO takes the low-latency tail kernel, the one with a chain the
            # general one: same arithmetic, different summation order)
            assert len(y) == len(y0) and (len(y) == 0 or rel_rms(y, y0) <= 1e-6)
            p = y
            if nb is not None:
                p = nb.process(p)
            if sq is not None:
                p = sq.process(p)
                muted_ref.append(not p.any())
                muted_gpu.append(fe.squelch_state(vid)[0])
            if fm is not None:
                p = fm.process(p)
            got.append(d); want.append(fe_ref(p)); raw.append(d0)
    got_blocks = got
    got, want, raw = np.concatenate(got), np.concatenate(want), np.concatenate(raw)
    assert len(got) == len(want)
    if fm is not None:
        # FMIF keeps the strongest DFT bin: where two bins are equally strong to fp32 rounding the GPU (fp32 DFT) and
        # the oracle (fp64 DFT) may pick different ones for that sample. Allow a handful of such samples, gate the rest.
        bad = np.abs(got - want) > 1e-2 * max(np.sqrt(np.mean(want.astype(np.float64) ** 2)), 1e-30)
        assert bad.sum() <= max(3, len(got) // 2000), int(bad.sum())
        got, want, raw = got[~bad], want[~bad], raw[~bad]
    assert rel_rms(got, want) <= tol, rel_rms(got, want)
    assert rel_rms(raw, want) > 100 * tol  # the chain did change the demodulator's input
    if sq is not None:
        assert muted_gpu == muted_ref
        assert any(muted_ref) and not all(muted_ref)
        # muted blocks are exactly zero
        for d, mu in zip(got_blocks, muted_ref):
            if mu and demod == po.DEMOD_AM:
                assert not d.any()


def test_if_chain_keeps_state_across_reconfiguration(gpu, port):
    """setLevel / setRate / enableBlock keep the blocks' state (noise_blanker.h:19-30, squelch.h:28-32)."""
    vfo = (24e3, 12e3, 100e3, po.DEMOD_AM)
    blocks = stream(30, vfo[2], 32)
    nb = port.noise_blanker(500.0 / 24e3, 1.6)
    got, want = [], []
    with gpu.Frontend(IN_SR, max_block=BLK) as fe:
        vid = fe.add_vfo(*vfo)
        fe.set_if_chain(vid, nb=True, nb_rate=500.0 / 24e3, nb_level=1.6)
        hist = []
        for i, b in enumerate(blocks):
            if i == 10:   # same object, new level: the running amplitude carries over
                fe.set_if_chain(vid, nb=True, nb_rate=500.0 / 24e3, nb_level=3.0)
                nb = _with_amp(port, 500.0 / 24e3, 3.0, hist)
            fe.process(po.FMT_CF32, b)
            y, d = fe.vfo_output(vid)
            hist.append(y)
            got.append(d); want.append(port.am_magnitude(nb.process(y)))
    assert rel_rms(np.concatenate(got), np.concatenate(want)) <= 1e-6


def _with_amp(port, rate, level, hist):
    """A NoiseBlanker at `level` whose running amplitude equals that of one at level 1.6 after `hist`: the amplitude
    recurrence does not depend on the level (noise_blanker.h:47-50), so replaying the history at any level gives it."""
    o = port.noise_blanker(rate, level)
    for y in hist:
        o.process(y)
    return o


def test_if_chain_errors(gpu):
    with gpu.Frontend(20e6, max_block=100000) as fe:
        v = fe.add_vfo(250e3, 200e3, 1e6, po.DEMOD_QUAD)
        fe.set_if_chain(v, squelch=True)        # 1250 samples per block: fits
        with pytest.raises(gpu.SdrppCudaError):
            fe.set_if_chain(v, fmif_bins=65)
        with pytest.raises(gpu.SdrppCudaError):
            fe.squelch_state(v + 5)
    with gpu.Frontend(2.4e6, max_block=1000000) as fe:
        v = fe.add_vfo(250e3, 200e3, 1e5, po.DEMOD_QUAD)
        with pytest.raises(gpu.SdrppCudaError):   # more outputs per block than the staging area holds
            fe.set_if_chain(v, squelch=True)
        with pytest.raises(gpu.SdrppCudaError):
            fe.squelch_state(v)


def test_if_chain_ragged_blocks(gpu, port):
    """Blocks of 12000, 4001, 7, 11995, 513 ... samples: VFO blocks of 0 .. 240 outputs, shorter than the FMIF window in
    places; the chain's state (blanker amplitude, squelch counter, FMIF history, previous sample) carries across."""
    vfo = (48e3, 12.5e3, 100e3, po.DEMOD_QUAD)
    x = np.concatenate(stream(12, vfo[2], 33))
    sizes = [12000, 4001, 7, 11995, 513, 12000, 1, 299, 12000] * 3
    sizes = sizes[:next(i for i in range(len(sizes)) if sum(sizes[:i + 1]) > len(x))]
    nb, sq, fm = port.noise_blanker(500.0 / 48e3, 1.6), port.squelch(-30.0), port.fm_if(31)
    q = port.quadrature(vfo[1] / 2.0, vfo[0])
    got, want, p0 = [], [], 0
    with gpu.Frontend(IN_SR, max_block=BLK) as fe:
        vid = fe.add_vfo(*vfo)
        fe.set_if_chain(vid, nb=True, nb_rate=500.0 / 48e3, nb_level=1.6, squelch=True, squelch_level=-30.0, fmif_bins=31)
        for s in sizes:
            fe.process(po.FMT_CF32, x[p0:p0 + s]); p0 += s
            y, d = fe.vfo_output(vid)
            assert len(d) == len(y)
            if len(y) == 0:
                continue   # the reference's blocks are never run on an empty block (rx_vfo.h:104-109: no swap)
            got.append(d); want.append(q.process(fm.process(sq.process(nb.process(y)))))
    assert min(len(g) for g in got) < 31 <= max(len(g) for g in got)
    got, want = np.concatenate(got), np.concatenate(want)
    bad = np.abs(got - want) > 1e-2 * np.sqrt(np.mean(want.astype(np.float64) ** 2))
    assert bad.sum() <= 3, int(bad.sum())
    assert rel_rms(got[~bad], want[~bad]) <= 2e-4
