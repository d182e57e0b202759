"""GPU: the command-list engine (csrc/launcher.h). A block is planned into a command list and replayed as instantiated CUDA
graphs once its command sequence has been seen twice; where the spectrum kernels run, whether cf32 ingest is fused into the
fp16 split and which tail kernel runs are scheduling choices. None of them may change a single bit of any result: the same
60-block stream through every variant, compared bitwise with the command-by-command run, and the graph statistics show that
the graph variants really replayed graphs (a suite of short tests alone would never reach the second sighting of a
sequence)."""
import os

import numpy as np
import pytest

from oracle import pyoracle as po
from sdrpp_b200 import synth

pytestmark = pytest.mark.gpu

SR, BLK, NB, FFT = 15.36e6, 76800, 60, 65536
VFOS = [(48e3, 12.5e3, -2.1e6, po.DEMOD_QUAD), (24e3, 12e3, 1.3e6, po.DEMOD_AM), (48e3, 2.7e3, 3.3e6, po.DEMOD_USB),
        (48e3, 12.5e3, 0.4e6, po.DEMOD_QUAD), (24e3, 12e3, -4.4e6, po.DEMOD_AM), (250e3, 200e3, 5.1e6, po.DEMOD_QUAD)]


def _stream():
    x = synth.baseband(NB * BLK, SR, 41, carriers=[(v[2], "fm") for v in VFOS[:4]], noise_dbfs=-45.0).astype(np.complex64)
    return [x[i * BLK:(i + 1) * BLK] for i in range(NB)]


def _run(gpu, blocks, env, ragged=False):
    keys = ("SDRPP_GRAPHS", "SDRPP_FFT_ORDER", "SDRPP_FUSE_INGEST", "SDRPP_TAIL_MODE")
    old = {k: os.environ.get(k) for k in keys}
    for k in keys:
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        out = {"iq": [[] for _ in VFOS], "dm": [[] for _ in VFOS], "rows": [], "rds": [], "audio": []}
        with gpu.Frontend(SR, fft_size=FFT, fft_rate=SR / FFT, fft_window=gpu.WIN_BH4, max_block=BLK) as fe:
            ids = [fe.add_vfo(*v) for v in VFOS]
            fe.set_post(ids[5], fm_lowpass=True, wfm=True, wfm_stereo=True, wfm_rds=True)
            for i, b in enumerate(blocks):
                if ragged and i % 7 == 3:
                    b = b[:BLK - 1 - i]
                fe.process(po.FMT_CF32, b)
                for k, vid in enumerate(ids):
                    y, d = fe.vfo_output(vid)
                    out["iq"][k].append(y); out["dm"][k].append(d)
                out["rows"].append(fe.fft_rows())
                out["rds"].append(fe.vfo_rds(ids[5]))
                out["audio"].append(np.stack(fe.vfo_audio_stereo(ids[5]), axis=1))
            stats = fe.graph_stats()
            tensor = fe.stage1_tensor_launches
        return out, stats, tensor
    finally:
        for k in keys:
            os.environ.pop(k, None)
            if old[k] is not None:
                os.environ[k] = old[k]


def _same(a, b):
    for k in range(len(VFOS)):
        for x, y in zip(a["iq"][k], b["iq"][k]):
            assert x.shape == y.shape and np.array_equal(x.view(np.uint32), y.view(np.uint32))
        for x, y in zip(a["dm"][k], b["dm"][k]):
            assert (x is None) == (y is None)
            if x is not None:
                assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
    for key in ("rows", "rds", "audio"):
        for x, y in zip(a[key], b[key]):
            assert x.shape == y.shape and np.array_equal(np.ascontiguousarray(x).view(np.uint32), np.ascontiguousarray(y).view(np.uint32)), key


@pytest.mark.parametrize("ragged", [False, True])
def test_graph_replay_and_scheduling_knobs_are_bit_identical(gpu, ragged):
    blocks = _stream()
    ref, st0, tensor = _run(gpu, blocks, {"SDRPP_GRAPHS": "0", "SDRPP_FFT_ORDER": "2", "SDRPP_FUSE_INGEST": "0"}, ragged)
    assert st0["replayed_runs"] == 0 and st0["graphs_instantiated"] == 0 and st0["direct_runs"] > 0
    assert tensor > 0                                    # the tensor-core stage 1 (and with it the fp16 split) is in play
    assert sum(len(r) for r in ref["rows"]) >= 60 and sum(len(r) for r in ref["rds"]) > 20
    variants = [{}, {"SDRPP_FFT_ORDER": "0"}, {"SDRPP_FFT_ORDER": "1"}, {"SDRPP_FFT_ORDER": "2"},
                {"SDRPP_FFT_ORDER": "0", "SDRPP_FUSE_INGEST": "0"}, {"SDRPP_GRAPHS": "0"}, {"SDRPP_TAIL_MODE": "general"},
                {"SDRPP_TAIL_MODE": "fast"}]
    for env in variants:
        got, st, _ = _run(gpu, blocks, env, ragged)
        if env.get("SDRPP_GRAPHS") != "0" and not ragged:
            assert st["replayed_runs"] > 5 and st["graphs_instantiated"] > 0, (env, st)
        if "SDRPP_TAIL_MODE" in env:
            # the two tail kernels add in a different order: same stream, <= 1e-6, counts identical
            for k in range(len(VFOS)):
                a, b = np.concatenate(got["iq"][k]), np.concatenate(ref["iq"][k])
                assert a.shape == b.shape
                err = np.sqrt(np.sum(np.abs(a.astype(np.complex128) - b) ** 2) / max(np.sum(np.abs(b.astype(np.complex128)) ** 2), 1e-30))
                assert err <= 1e-6, (env, k, err)
            for x, y in zip(got["rows"], ref["rows"]):
                assert np.array_equal(x.view(np.uint32), y.view(np.uint32))
        else:
            _same(got, ref)


def test_control_calls_between_replayed_blocks(gpu):
    """Retune, add and remove a VFO in the middle of a stream that is being replayed as graphs: the layout change makes new
    command sequences (new graphs), results stay those of the command-by-command run."""
    blocks = _stream()[:40]

    def run(env):
        old = os.environ.get("SDRPP_GRAPHS")
        os.environ.update(env)
        try:
            res = []
            with gpu.Frontend(SR, fft_size=FFT, fft_rate=SR / FFT, fft_window=gpu.WIN_BH4, max_block=BLK) as fe:
                ids = [fe.add_vfo(*v) for v in VFOS[:4]]
                for i, b in enumerate(blocks):
                    if i == 14:
                        fe.vfo_set_offset(ids[0], -2.0e6)
                    if i == 20:
                        ids.append(fe.add_vfo(*VFOS[4]))
                    if i == 30:
                        fe.remove_vfo(ids.pop(1))
                    fe.process(po.FMT_CF32, b)
                    res.append([fe.vfo_output(v)[0] for v in ids] + [fe.fft_rows()])
                st = fe.graph_stats()
            return res, st
        finally:
            os.environ.pop("SDRPP_GRAPHS", None)
            if old is not None:
                os.environ["SDRPP_GRAPHS"] = old

    a, sa = run({"SDRPP_GRAPHS": "1"})
    b, sb = run({"SDRPP_GRAPHS": "0"})
    assert sa["replayed_runs"] > 0 and sb["replayed_runs"] == 0
    for ra, rb in zip(a, b):
        assert len(ra) == len(rb)
        for x, y in zip(ra, rb):
            assert x.shape == y.shape and np.array_equal(np.ascontiguousarray(x).view(np.uint32), np.ascontiguousarray(y).view(np.uint32))
