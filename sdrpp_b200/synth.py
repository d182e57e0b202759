"""Synthetic IQ shared by the parity tests and bench.py: tones on a rational grid,
FM/AM carriers at the VFO centres and white noise, generated in fp64 from PCG64(seed), then
quantised to the configured sample format."""
import numpy as np

FMT_CF32, FMT_U8_RTL, FMT_U8_TCP, FMT_I8, FMT_I16_FILE, FMT_I16_VOLK = range(6)


def baseband(n, fs, seed, carriers=(), noise_dbfs=-70.0, tones=6, n0=0):
    """n complex samples starting at absolute index n0. carriers: list of (offset_hz, kind) with kind in
    {'fm','am','cw'}."""
    rng = np.random.Generator(np.random.PCG64(seed))
    t = (np.arange(n, dtype=np.float64) + n0) / fs
    x = np.zeros(n, dtype=np.complex128)
    for i in range(tones):
        f = fs * (-0.45 + 0.9 * ((7 * i + 3) % 23) / 23.0)
        a = 10.0 ** (-(6.0 + 9.0 * i) / 20.0)
        x += a * np.exp(2j * np.pi * (f * t + 0.1 * i))
    for j, (off, kind) in enumerate(carriers):
        a = 10.0 ** (-20.0 / 20.0)
        fm = 400.0 + 50.0 * (j % 7)
        if kind == "fm":
            dev = 3000.0
            x += a * np.exp(2j * np.pi * off * t + 1j * (dev / fm) * np.sin(2 * np.pi * fm * t))
        elif kind == "am":
            x += a * (1.0 + 0.5 * np.sin(2 * np.pi * fm * t)) * np.exp(2j * np.pi * off * t)
        else:
            x += a * np.exp(2j * np.pi * off * t)
    sigma = 10.0 ** (noise_dbfs / 20.0) / np.sqrt(2.0)
    x += sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    peak = np.max(np.abs(np.concatenate([x.real, x.imag])))
    if peak > 0.95:
        x *= 0.95 / peak
    return x


def quantise(x, fmt):
    """complex128 -> raw array of the given format (interleaved I,Q for integer formats)."""
    if fmt == FMT_CF32:
        return x.astype(np.complex64)
    iq = np.empty(2 * len(x), dtype=np.float64)
    iq[0::2] = x.real
    iq[1::2] = x.imag
    if fmt in (FMT_U8_RTL, FMT_U8_TCP):
        return np.clip(np.round(127.5 * iq + 127.5), 0, 255).astype(np.uint8)
    if fmt == FMT_I8:
        return np.clip(np.round(127.5 * iq - 0.5), -128, 127).astype(np.int8)
    return np.clip(np.round(32767.5 * iq - 0.5), -32768, 32767).astype(np.int16)


def vfo_grid(nvfo, fs, span=0.9):
    """VFO centre offsets on a uniform grid inside +-span/2 * fs."""
    if nvfo == 1:
        return np.array([0.1 * fs / 2.4])
    return (np.arange(nvfo) - (nvfo - 1) / 2.0) * (span * fs / nvfo)
