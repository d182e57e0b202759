"""GPU parity of the tensor-core stage 1 (channelizer_tc.cu): FrequencyXlator + first DecimatingFIR as a
tcgen05 matrix product with fp16 hi/lo split operands.

Checked three ways, all through the C ABI: (1) against the oracle with offset 0 (NCO = identity) at the
1e-5 gate of SURVEY 8d; (2) against the library's own FP32 FMA stage 1 (mode 1) on identical inputs -- both
evaluate the same ideal NCO, so they must agree far inside the gate, whatever the VFO offsets; (3) with the
NCO against the oracle's ideal-NCO flavour, directly, at 1e-5. Every test asserts that the
tensor-core kernel actually ran."""
import numpy as np
import pytest

from oracle import pyoracle as po
from sdrpp_b200 import synth

pytestmark = pytest.mark.gpu

# (inSR, outSR, bw, block): PowerDecimator ratios 256, 512 (first stage /32) and 1024, 2048, 4096 (first stage /64)
TC_PLANS = [
    (15.36e6, 48e3, 2.7e3, 76800),
    (15.36e6, 24e3, 12e3, 76800),
    (61.44e6, 48e3, 12.5e3, 307200),
    (122.88e6, 48e3, 12.5e3, 614400),
    (122.88e6, 24e3, 12e3, 614400),
]


def run(gpu, sr, vfos, blocks, mode, max_block=None, retune=None, **kw):
    """Per VFO the concatenated cf32 output and the per-block counts; retune = (block index, vfo index, offset)."""
    mb = max_block or max(len(b) for b in blocks)
    out = [[] for _ in vfos]
    with gpu.Frontend(sr, max_block=mb, **kw) as fe:
        fe.set_stage1_mode(mode)
        ids = [fe.add_vfo(*v) for v in vfos]
        for bi, b in enumerate(blocks):
            if retune is not None and retune[0] == bi:
                fe.vfo_set_offset(ids[retune[1]], retune[2])
            fe.process(po.FMT_CF32, b)
            for i, vid in enumerate(ids):
                out[i].append(fe.vfo_output(vid)[0].copy())
        tl = fe.stage1_tensor_launches
    return [np.concatenate(o) for o in out], [[len(a) for a in o] for o in out], tl


@pytest.mark.parametrize("inSR,outSR,bw,blk", TC_PLANS)
def test_offset0_against_oracle(gpu, port, inSR, outSR, bw, blk):
    nblocks = 4
    x = synth.baseband(blk * nblocks, inSR, 21, carriers=[(0.0, "fm"), (bw, "am")], noise_dbfs=-40.0).astype(np.complex64)
    blocks = [x[i * blk:(i + 1) * blk] for i in range(nblocks)]
    g, counts, tl = run(gpu, inSR, [(outSR, bw, 0.0, po.DEMOD_NONE)], blocks, 0)
    assert tl >= nblocks - 1, "tensor-core stage 1 did not run"
    o = port.rxvfo(inSR, outSR, bw, 0.0)
    r = [o.process(b) for b in blocks]
    assert counts[0] == [len(a) for a in r]
    err = po.rel_rms(g[0], np.concatenate(r))
    assert err <= 1e-5, f"rel-RMS {err:.3e}"


@pytest.mark.parametrize("inSR,outSR,bw,blk", TC_PLANS)
def test_matches_fp32_kernel_many_vfos(gpu, inSR, outSR, bw, blk):
    """37 VFOs (three tiles, the last one partial) at arbitrary offsets, ragged blocks."""
    sizes = [blk, blk // 3 + 1, 7, blk - 5, 513, blk]
    offs = synth.vfo_grid(37, inSR)
    x = synth.baseband(sum(sizes), inSR, 22, carriers=[(float(o), "fm") for o in offs[::4]], noise_dbfs=-50.0).astype(np.complex64)
    blocks, p = [], 0
    for s in sizes:
        blocks.append(x[p:p + s]); p += s
    vfos = [(outSR, bw, float(o), po.DEMOD_NONE) for o in offs]
    gt, ct, tl = run(gpu, inSR, vfos, blocks, 0, max_block=blk)
    gf, cf, tl0 = run(gpu, inSR, vfos, blocks, 1, max_block=blk)
    assert tl > 0 and tl0 == 0
    assert ct == cf
    x_rms = float(np.sqrt(np.mean(np.abs(x) ** 2)))
    worst = 0.0
    for i in range(len(vfos)):
        ref_rms = float(np.sqrt(np.mean(np.abs(gf[i]) ** 2)))
        d = float(np.sqrt(np.mean(np.abs(gt[i] - gf[i]) ** 2)))
        # channels without a carrier hold only noise: absolute gate relative to the full-band signal there
        assert d <= 3e-6 * ref_rms + 1e-7 * x_rms, f"vfo {i}: diff {d:.3e} ref {ref_rms:.3e}"
        worst = max(worst, d / max(ref_rms, 1e-12))
    print(f"worst tensor-vs-fp32 relative difference {worst:.3e}")


@pytest.mark.parametrize("inSR,outSR,bw,blk", TC_PLANS)
def test_nco_against_ideal_oracle(gpu, port, report, inSR, outSR, bw, blk):
    """With the NCO, against the oracle's ideal-NCO flavour: direct, per block, <= 1e-5 (the reference rotator's own walk
    away from the same ideal is reported beside it)."""
    off = 0.2137 * inSR / 2.4
    nblocks = 3
    x = synth.baseband(blk * nblocks, inSR, 23, carriers=[(off, "fm")], noise_dbfs=-40.0).astype(np.complex64)
    blocks = [x[i * blk:(i + 1) * blk] for i in range(nblocks)]
    g, counts, tl = run(gpu, inSR, [(outSR, bw, off, po.DEMOD_NONE)], blocks, 0)
    assert tl >= nblocks - 1
    o, o32 = port.rxvfo(inSR, outSR, bw, off, ideal_nco=True), port.rxvfo(inSR, outSR, bw, off)
    p = 0
    worst, walk = 0.0, 0.0
    for b in range(nblocks):
        r, r32 = o.process(blocks[b]), o32.process(blocks[b])
        assert counts[0][b] == len(r)
        err = po.rel_rms(g[0][p:p + len(r)], r)
        worst, walk = max(worst, err), max(walk, po.rel_rms(r32, r))
        p += len(r)
        assert err <= 1e-5, f"block {b}: {err:.3e}"
    report(f"A11 tensor stage 1 {inSR/1e6:g}M->{outSR/1e3:g}k", gpu_vs_ideal_worst_block=worst, ref_f32_vs_ideal_worst_block=walk, gate=1e-5)


@pytest.mark.parametrize("scale", [1e-6, 1.0, 3000.0])
def test_block_scaling_and_ring_wrap(gpu, scale):
    """Input amplitude far from full scale (the fp16 split is block-scaled per 8 rows) on a ring that wraps
    several times during the run."""
    inSR, outSR, bw, blk = 15.36e6, 24e3, 12e3, 76800
    nblocks = 9
    offs = synth.vfo_grid(5, inSR)
    x = (synth.baseband(blk * nblocks, inSR, 24, carriers=[(float(o), "am") for o in offs], noise_dbfs=-50.0) * scale).astype(np.complex64)
    # a quiet stretch and a loud burst inside one block: neighbouring row groups get different exponents
    x[blk * 4 + 1000: blk * 4 + 20000] *= 1e-3
    x[blk * 5 + 7: blk * 5 + 3000] *= 30.0
    blocks = [x[i * blk:(i + 1) * blk] for i in range(nblocks)]
    vfos = [(outSR, bw, float(o), po.DEMOD_NONE) for o in offs]
    gt, ct, tl = run(gpu, inSR, vfos, blocks, 0, ring_log2=18)
    gf, cf, _ = run(gpu, inSR, vfos, blocks, 1, ring_log2=18)
    assert tl > 0 and ct == cf
    for i in range(len(vfos)):
        err = po.rel_rms(gt[i], gf[i])
        assert err <= 3e-6, f"vfo {i}: {err:.3e}"


def test_retune_midstream(gpu):
    inSR, outSR, bw, blk = 61.44e6, 48e3, 12.5e3, 307200
    nblocks = 5
    x = synth.baseband(blk * nblocks, inSR, 25, carriers=[(1.0e6, "fm"), (-7.3e6, "fm")], noise_dbfs=-40.0).astype(np.complex64)
    blocks = [x[i * blk:(i + 1) * blk] for i in range(nblocks)]
    vfos = [(outSR, bw, 1.0e6, po.DEMOD_NONE), (outSR, bw, 3.0e6, po.DEMOD_NONE)]
    gt, ct, tl = run(gpu, inSR, vfos, blocks, 0, retune=(2, 1, -7.3e6))
    gf, cf, _ = run(gpu, inSR, vfos, blocks, 1, retune=(2, 1, -7.3e6))
    assert tl > 0 and ct == cf
    for i in range(2):
        err = po.rel_rms(gt[i], gf[i])
        assert err <= 3e-6, f"vfo {i}: {err:.3e}"


def _compare_modes(gpu, sr, blocks, script, max_block, tol=3e-6):
    """script(fe, block_index) -> may add / retune VFOs before the block; returns the list of live VFO ids. Runs the same
    script in tensor and FP32 mode and compares every VFO's concatenated output."""
    outs = []
    for mode in (0, 1):
        res = {}
        with gpu.Frontend(sr, max_block=max_block) as fe:
            fe.set_stage1_mode(mode)
            state = {}
            for bi, b in enumerate(blocks):
                ids = script(fe, bi, state)
                fe.process(po.FMT_CF32, b)
                for name, vid in ids.items():
                    res.setdefault(name, []).append(fe.vfo_output(vid)[0].copy())
            tl = fe.stage1_tensor_launches
        outs.append(({k: np.concatenate(v) for k, v in res.items()}, tl))
    (gt, tlt), (gf, tlf) = outs
    assert tlt > 0 and tlf == 0
    assert gt.keys() == gf.keys()
    x_rms = float(np.sqrt(np.mean(np.abs(np.concatenate(blocks)) ** 2)))
    for k in gt:
        assert len(gt[k]) == len(gf[k]), k
        ref_rms = float(np.sqrt(np.mean(np.abs(gf[k]) ** 2)))
        d = float(np.sqrt(np.mean(np.abs(gt[k] - gf[k]) ** 2)))
        assert d <= tol * ref_rms + 1e-7 * x_rms, f"{k}: diff {d:.3e} ref {ref_rms:.3e}"


def test_both_plane_sets_at_once(gpu):
    """First-stage decimations 32 and 64 in the same front end: two sets of fp16 planes, two launches per block."""
    sr, blk = 15.36e6, 76800
    x = synth.baseband(blk * 5, sr, 31, carriers=[(1.0e6, "am"), (-2.0e6, "am")], noise_dbfs=-40.0).astype(np.complex64)
    blocks = [x[i * blk:(i + 1) * blk] for i in range(5)]

    def script(fe, bi, st):
        if bi == 0:
            st["a"] = fe.add_vfo(24e3, 12e3, 1.0e6, po.DEMOD_NONE)     # ratio 512: first stage /32
            st["b"] = fe.add_vfo(12e3, 6e3, -2.0e6, po.DEMOD_NONE)     # ratio 1024: first stage /64
            assert fe.vfo_info(st["a"])["s1_decim"] == 32 and fe.vfo_info(st["b"])["s1_decim"] == 64
        return dict(st)

    _compare_modes(gpu, sr, blocks, script, blk)


def test_more_groups_than_one_launch_holds(gpu):
    """Eight plans of the same first stage = eight groups: the tensor kernel is launched twice per block (6 + 2)."""
    sr, blk = 122.88e6, 614400
    bws = [6e3, 7e3, 8e3, 9e3, 10e3, 11e3, 12e3, 12.5e3]
    offs = [float(o) for o in synth.vfo_grid(len(bws), sr)]
    x = synth.baseband(blk * 3, sr, 32, carriers=[(o, "fm") for o in offs], noise_dbfs=-50.0).astype(np.complex64)
    blocks = [x[i * blk:(i + 1) * blk] for i in range(3)]

    def script(fe, bi, st):
        if bi == 0:
            for i, (bw, o) in enumerate(zip(bws, offs)):
                st[f"v{i}"] = fe.add_vfo(48e3, bw, o, po.DEMOD_NONE)
        return dict(st)

    _compare_modes(gpu, sr, blocks, script, blk)


def test_vfo_class_added_midstream_moves_the_row_origin(gpu):
    """A second class of VFOs appears after two blocks: the row origin is chosen again for both classes and the history
    the next windows reach back into is converted again; the first class must not notice."""
    sr, blk = 122.88e6, 614400
    x = synth.baseband(blk * 6, sr, 33, carriers=[(5.0e6, "fm"), (-11.0e6, "am")], noise_dbfs=-40.0).astype(np.complex64)
    blocks = [x[i * blk:(i + 1) * blk] for i in range(6)]

    def script(fe, bi, st):
        if bi == 0:
            st["nfm"] = fe.add_vfo(48e3, 12.5e3, 5.0e6, po.DEMOD_NONE)
        if bi == 2:
            for i in range(40):                                        # enough AM VFOs to outweigh the single NFM one
                st[f"am{i}"] = fe.add_vfo(24e3, 12e3, -11.0e6 + 25e3 * i, po.DEMOD_NONE)
        return dict(st)

    _compare_modes(gpu, sr, blocks, script, blk)


def test_mode_switch_on_a_live_front_end(gpu):
    """Tensor -> FP32 -> tensor on a live front end: while the FP32 kernel runs, the fp16 planes are not refreshed, so the
    first tensor-core block after switching back must not read them for history (it falls back to the FP32 kernel for the
    blocks whose windows reach before the planes' valid range)."""
    sr, blk = 61.44e6, 307200
    x = synth.baseband(blk * 8, sr, 35, carriers=[(2.0e6, "fm")], noise_dbfs=-40.0).astype(np.complex64)
    blocks = [x[i * blk:(i + 1) * blk] for i in range(8)]
    outs = []
    for toggled in (True, False):
        res = []
        with gpu.Frontend(sr, max_block=blk) as fe:
            fe.set_stage1_mode(0 if toggled else 1)
            vid = fe.add_vfo(48e3, 12.5e3, 2.0e6, po.DEMOD_NONE)
            for bi, b in enumerate(blocks):
                if toggled and bi == 3:
                    fe.set_stage1_mode(1)
                if toggled and bi == 5:
                    fe.set_stage1_mode(0)
                fe.process(po.FMT_CF32, b)
                res.append(fe.vfo_output(vid)[0].copy())
            tl = fe.stage1_tensor_launches
        outs.append((np.concatenate(res), tl))
    (gt, tlt), (gf, tlf) = outs
    assert tlt >= 3 and tlf == 0
    assert len(gt) == len(gf)
    err = po.rel_rms(gt, gf)
    assert err <= 3e-6, f"{err:.3e}"
