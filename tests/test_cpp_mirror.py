"""The C++ drop-in (include/sdrpp/ + sdrpp_b200/host/): dsp::stream, dsp::channel::RxVFO and IQFrontEnd replaced over
the C ABI, everything else -- dsp::block, dsp::Processor, dsp::complex_t, threading, logging -- the reference's own
files through the source overlay. A small GUI-less C++ program (tests/cpp/mirror_demo.cpp) uses them exactly as an
SDR++ module would and is checked against the oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pyoracle as po
from sdrpp_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "tests", "cpp", "mirror_demo")


@pytest.fixture(scope="module")
def demo(cuda_lib):
    """The demo is compiled against the source overlay of the reference tree (tools/make_overlay.py), so it can only be
    (re)built where /root/reference exists; on the GPU box the binary built here is used if its content stamp is current."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_overlay
    if os.path.isdir("/root/reference/core/src"):
        return make_overlay.build_demo()
    if not make_overlay.demo_is_current():
        pytest.fail("tests/cpp/mirror_demo is missing or stale and the reference tree is not here to rebuild it: "
                    "run __graft_entry__.build() in the build container")
    return make_overlay.DEMO


def test_host_side_stream_block_semantics(demo):
    """No GPU: blocking swap/read/flush, idempotent start/stop, nested tempStop, stop flags, waterfall offset maths."""
    r = subprocess.run([demo, "host"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_module_style_usage_matches_oracle(demo, gpu, port, tmp_path):
    sr, blk, N, outSR, bw = 2.4e6, 12000, 8192, 240e3, 200e3
    offs = [100e3, -300e3]
    nblocks = 8
    x = synth.baseband(blk * nblocks, sr, 31, carriers=[(o, "fm") for o in offs], noise_dbfs=-40.0).astype(np.complex64)
    inp = tmp_path / "in.cf32"
    x.tofile(inp)
    prefix = str(tmp_path / "out")
    r = subprocess.run([demo, "run", str(inp), str(sr), str(blk), str(N), prefix, str(outSR), str(bw)] + [str(o) for o in offs],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    for i, off in enumerate(offs):
        y = np.fromfile(prefix + f".vfo{i}.cf32", dtype=np.complex64)
        v = port.rxvfo(sr, outSR, bw, off, ideal_nco=True)   # the oracle's ideal-NCO flavour
        ref = np.concatenate([v.process(x[b * blk:(b + 1) * blk]) for b in range(nblocks)])
        assert len(y) == len(ref)
        for b in range(nblocks):
            seg = slice(b * len(ref) // nblocks, (b + 1) * len(ref) // nblocks)
            res = po.rel_rms(y[seg], ref[seg])
            assert res <= 1e-5, (i, b, res)
    solo = np.fromfile(prefix + ".solo.cf32", dtype=np.complex64)
    y0 = np.fromfile(prefix + ".vfo0.cf32", dtype=np.complex64)
    assert len(solo) == len(y0) and po.rel_rms(solo, y0) <= 1e-6  # standalone RxVFO == attached VFO
    rows = np.fromfile(prefix + ".rows.f32", dtype=np.float32).reshape(-1, N)
    assert rows.shape[0] == (blk * nblocks) // N
    w = port.window(po.WIN_BH7, N)
    for f in (0, rows.shape[0] - 1):
        _, _, row64 = port.spectrum(N, x[f * N:(f + 1) * N], w)
        mask = row64 >= row64.max() - 80.0
        assert np.abs(rows[f] - row64)[mask].max() <= 0.01


@pytest.mark.gpu
@pytest.mark.parametrize("ptype", [0, 1, 2])
def test_compression_blocks_match_oracle(demo, gpu, port, tmp_path, ptype):
    """dsp::compression::SampleStreamCompressor::process / SampleStreamDecompressor::process of the mirror, called the
    way core/src/server.cpp and the sdrpp_server source call them: packet and samples bit-identical to the oracle."""
    from tools.make_golden import pcm_input
    x = pcm_input(30001, 50 + ptype)
    inp = tmp_path / "in.cf32"
    x.tofile(inp)
    pk, out = tmp_path / "pk.bin", tmp_path / "out.cf32"
    r = subprocess.run([demo, "pcm", str(inp), str(ptype), str(pk), str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    packet = np.fromfile(pk, dtype=np.uint8)
    want = port.pcm_compress(ptype, x)
    assert np.array_equal(packet, want)
    y = np.fromfile(out, dtype=np.complex64)
    assert np.array_equal(y.view(np.uint32), port.pcm_decompress(want).view(np.uint32))
