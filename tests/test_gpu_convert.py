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
