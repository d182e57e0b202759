"""GPU parity: source sample conversions -- bit-exact against the oracle."""
import numpy as np
import pytest

from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

FMTS8 = [po.FMT_U8_RTL, po.FMT_U8_TCP, po.FMT_I8]
FMTS16 = [po.FMT_I16_FILE, po.FMT_I16_VOLK]


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("fmt", FMTS8)
def test_full_sweep_8bit(gpu, port, fmt):
    vals = np.arange(256, dtype=np.uint8)
    raw = np.repeat(vals, 2)  # I = Q = v
    if fmt == po.FMT_I8:
        raw = raw.view(np.int8)
    out = gpu.convert(fmt, raw)
    assert np.array_equal(_bits(out), _bits(port.convert(fmt, raw)))


@pytest.mark.parametrize("fmt", FMTS16)
def test_full_sweep_16bit(gpu, port, fmt):
    raw = np.repeat(np.arange(-32768, 32768, dtype=np.int32).astype(np.int16), 2)
    out = gpu.convert(fmt, raw)
    assert np.array_equal(_bits(out), _bits(port.convert(fmt, raw)))


def test_known_answer_hashes(gpu):
    # word-wise FNV-1a of the IEEE tables
    t8 = gpu.convert(po.FMT_U8_RTL, np.repeat(np.arange(256, dtype=np.uint8), 2)).view(np.float32)[0::2]
    assert po.fnv1a_words(t8) == 0xb8121cc5
    assert t8[0] == np.float32(-1.0) and t8[255] == np.float32(1.0) and t8[128] == np.float32(0.00392156886)
    t16 = gpu.convert(po.FMT_I16_FILE, np.repeat(np.arange(-32768, 32768, dtype=np.int32).astype(np.int16), 2)).view(np.float32)[0::2]
    assert po.fnv1a_words(t16) == 0x4fdefa45


@pytest.mark.parametrize("fmt", [po.FMT_CF32] + FMTS8 + FMTS16)
@pytest.mark.parametrize("n", [1, 2, 3, 5, 7936, 100003])
def test_random_blocks(gpu, port, fmt, n):
    rng = np.random.default_rng(100 * fmt + n)
    if fmt == po.FMT_CF32:
        raw = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    elif fmt == po.FMT_I8:
        raw = rng.integers(-128, 128, 2 * n).astype(np.int8)
    elif fmt in FMTS8:
        raw = rng.integers(0, 256, 2 * n).astype(np.uint8)
    else:
        raw = rng.integers(-32768, 32768, 2 * n).astype(np.int16)
    out = gpu.convert(fmt, raw)
    assert np.array_equal(_bits(out), _bits(port.convert(fmt, raw)))


def _raw_wide(fmt, n, rng):
    if fmt == po.FMT_I24_FILE:
        v = rng.integers(-(1 << 23), 1 << 23, 2 * n).astype(np.int32)
        v[:4] = [-(1 << 23), (1 << 23) - 1, 0, -1]
        return np.stack([(v & 255), (v >> 8) & 255, (v >> 16) & 255], axis=1).astype(np.uint8).reshape(-1)
    if fmt == po.FMT_I32_FILE:
        v = rng.integers(-(1 << 31), 1 << 31, 2 * n).astype(np.int32)
        v[:4] = [-(1 << 31), (1 << 31) - 1, 0, -1]
        return v
    v = rng.standard_normal(2 * n) * 10.0 ** rng.uniform(-8, 2, 2 * n)
    v[:4] = [0.0, -0.0, 1.0 + 2.0 ** -24, 3.4e38]
    return v.astype(np.float64)


@pytest.mark.parametrize("fmt", [po.FMT_I24_FILE, po.FMT_I32_FILE, po.FMT_F64])
@pytest.mark.parametrize("n", [2, 5, 7936, 100003])
def test_wide_file_formats(gpu, port, fmt, n):
    """WAV i24 / i32 / f64 (file_source/src/main.cpp:470-545): SURVEY 8f rank 3."""
    raw = _raw_wide(fmt, n, np.random.default_rng(7 * fmt + n))
    out = gpu.convert(fmt, raw)
    assert len(out) == n
    assert np.array_equal(_bits(out), _bits(port.convert(fmt, raw)))


def test_wide_format_through_frontend(gpu, port):
    n = 4096
    raw = _raw_wide(po.FMT_I24_FILE, n, np.random.default_rng(3))
    with gpu.Frontend(2.4e6, max_block=n) as fe:
        fe.process(po.FMT_I24_FILE, raw)
        iq = fe.read_iq(n)
    assert np.array_equal(_bits(iq), _bits(port.convert(po.FMT_I24_FILE, raw)))


def test_empty_and_bad_args(gpu):
    assert len(gpu.convert(po.FMT_U8_RTL, np.zeros(0, np.uint8))) == 0
    with pytest.raises(gpu.SdrppCudaError):
        gpu._check(gpu.lib().sdrpp_cuda_convert(99, None, 4, None), "convert")


# ---- SDR++ server wire packets (dsp/compression) ---------------------------------------------
def _pcm_block(n, seed, scale=0.3):
    rng = np.random.default_rng(seed)
    x = (scale * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
    if n >= 3:
        x[n // 3] = np.complex64(-1.7 * scale / 0.3 + 0.2j)  # below -max: the saturation branch
    return x


@pytest.mark.parametrize("ptype", [0, 1, 2])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 7936, 100003, 1000000])
def test_pcm_compress_bit_exact(gpu, port, ptype, n):
    """SampleStreamCompressor::process (sample_stream_compressor.h:26-60): header, scaler and payload bytes."""
    x = _pcm_block(n, 10 * n + ptype)
    got, want = gpu.pcm_compress(ptype, x), port.pcm_compress(ptype, x)
    assert len(got) == 8 + n * (2, 4, 8)[ptype]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("ptype", [0, 1, 2])
@pytest.mark.parametrize("n", [1, 2, 3, 5, 7936, 100003, 1000000])
def test_pcm_decompress_bit_exact(gpu, port, ptype, n):
    """SampleStreamDecompressor::process (sample_stream_decompressor.h:13-36) on packets with odd scalers."""
    pk = port.pcm_compress(ptype, _pcm_block(n, 20 * n + ptype, scale=0.0137))
    got, want = gpu.pcm_decompress(pk), port.pcm_decompress(pk)
    assert len(got) == n
    assert np.array_equal(_bits(got), _bits(want))


@pytest.mark.parametrize("ptype", [0, 1])
def test_pcm_full_code_sweep(gpu, port, ptype):
    """Every int8 / int16 code through the decompressor with a scaler that is not a power of two."""
    codes = np.arange(-128, 128, dtype=np.int8) if ptype == 0 else np.arange(-32768, 32768, dtype=np.int32).astype(np.int16)
    payload = np.repeat(codes, 2).view(np.uint8)
    for scaler in (0.7310586, 1.0, 3.0e-5):
        hdr = np.zeros(8, np.uint8)
        hdr[2:4] = np.array([ptype], np.uint16).view(np.uint8)
        hdr[4:8] = np.array([scaler], np.float32).view(np.uint8)
        pk = np.concatenate([hdr, payload])
        assert np.array_equal(_bits(gpu.pcm_decompress(pk)), _bits(port.pcm_decompress(pk)))


def test_pcm_golden(gpu):
    """Packets made and read back by the reference's own headers (tests/golden/pcm_*.npz)."""
    import json, os
    from tools.make_golden import pcm_input
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    cases = [c for c in json.load(open(os.path.join(gold, "manifest.json")))["cases"] if c["kind"] == "pcm"]
    assert len(cases) == 3
    for c in cases:
        d = np.load(os.path.join(gold, c["file"]))
        ptype, n = c["args"]
        assert np.array_equal(gpu.pcm_compress(ptype, pcm_input(n, c["seed"])), d["packet"])
        assert np.array_equal(_bits(gpu.pcm_decompress(d["packet"])), _bits(d["out"]))


def test_pcm_round_trip_error_bound(gpu):
    """Size-independent property: decompress(compress(x)) is within half a code step where x does not saturate."""
    x = _pcm_block(614400, 5)
    mx = x.view(np.float32).max()
    for ptype, full in ((0, 128.0), (1, 32768.0)):
        y = gpu.pcm_decompress(gpu.pcm_compress(ptype, x))
        a, b = x.view(np.float32).astype(np.float64), y.view(np.float32).astype(np.float64)
        ok = (a * (full / mx) <= full - 1) & (a * (full / mx) >= -full)
        assert ok.sum() > 0.99 * len(a)
        # half a code step plus the fp32 rounding of x*(full/mx) and code/(full/mx) at |x| < 2
        assert np.max(np.abs(a[ok] - b[ok])) <= 0.5 * mx / full + 5e-7


@pytest.mark.parametrize("ptype", [0, 1, 2])
def test_pcm_packet_through_frontend(gpu, port, ptype):
    """sdrpp_server_source: packets feed the front end; ring contents = the decompressor's output, across blocks."""
    n = 4099
    pks = [port.pcm_compress(ptype, _pcm_block(n, 30 + i + ptype)) for i in range(3)]
    with gpu.Frontend(2.4e6, max_block=n) as fe:
        for pk in pks:
            assert fe.submit_pcm(pk) == n
            fe.wait()
            assert np.array_equal(_bits(fe.read_iq(n)), _bits(port.pcm_decompress(pk)))


def test_pcm_unknown_type_and_bad_args(gpu):
    pk = np.zeros(24, np.uint8)
    pk[2] = 9
    assert len(gpu.pcm_decompress(pk)) == 0
    with gpu.Frontend(2.4e6, max_block=64) as fe:
        assert fe.submit_pcm(pk) == 0
    with pytest.raises(gpu.SdrppCudaError):
        gpu._check(gpu.lib().sdrpp_cuda_pcm_decompress(None, 4, None), "pcm")
    with pytest.raises(gpu.SdrppCudaError):
        gpu.pcm_compress(5, np.zeros(4, np.complex64))
