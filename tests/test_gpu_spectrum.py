"""GPU parity: fused window + FFT + dB spectrum line against the oracle.

Gates: complex X <= 1e-5 relative RMS vs the fp64 DFT of the fp32-windowed frame;
row <= 0.01 dB on bins within 100 dB of the line peak (App. C.10 explains the level mask)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from sdrpp_b200 import synth

pytestmark = pytest.mark.gpu

TOL_X = 1e-5
TOL_DB = 0.01
SCATTER_X = 4.0  # N >= 64K only (App. C.10): allowed ratio to the oracle's own fp32-FFT error on the gated bins


def _frame(n, seed, noise=-40.0):
    return synth.baseband(n, 2.4e6, seed, carriers=[(300e3, "fm"), (-500e3, "am")], noise_dbfs=noise).astype(np.complex64)


def _check(gpu, port, N, nz, wtype, seed, report=None):
    x = _frame(nz, seed)
    w = port.window(wtype, nz)
    row32, X64, row64 = port.spectrum(N, x, w)
    row, X = gpu.spectrum(N, x, w, want_X=True)
    err = po.rel_rms(X, X64)
    print(f"N={N} nz={nz} X rel-RMS {err:.3e}")
    assert err <= TOL_X, f"N={N} nz={nz}: X rel-RMS {err:.3e}"
    mask = row64 >= row64.max() - 100.0
    d = np.abs(row.astype(np.float64) - row64)[mask]
    # 0.01 dB on the gated bins, strictly, below 64K points. From 64K up the per-bin noise floor of the test signal comes
    # within 10 dB of the 100 dB mask, where a few bins sit at the fp32 round-off of ANY transform (the oracle's own
    # fp32 FFT reads 0.006 ... 0.03 dB there, App. C.10): there, no worse than SCATTER_X times the oracle's error
    d_ref = np.abs(row32.astype(np.float64) - row64)[mask]
    gate = max(TOL_DB, SCATTER_X * d_ref.max()) if N >= (1 << 16) else TOL_DB
    if report is not None:
        report(f"A10 spectrum N={N} nz={nz} win{wtype}", X_rel_rms=err, row_db_within_100dB=float(d.max()), oracle_fp32_fft_db=float(d_ref.max()), gate_db=float(gate))
    assert d.max() <= gate, f"N={N}: max |dB| {d.max():.4f} (ref fp32 {d_ref.max():.4f}) on {mask.sum()} gated bins"
    strong = row64 >= row64.max() - 60.0
    assert np.abs(row.astype(np.float64) - row64)[strong].max() <= TOL_DB
    # the GPU must not be worse than the reference's own fp32 scatter below the gate
    ref_scatter = np.percentile(np.abs(row32.astype(np.float64) - row64), 99.9)
    gpu_scatter = np.percentile(np.abs(row.astype(np.float64) - row64), 99.9)
    assert gpu_scatter <= max(4.0 * ref_scatter, TOL_DB), (gpu_scatter, ref_scatter)


@pytest.mark.parametrize("N", [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072, 262144, 524288, 1048576])
def test_sizes_bh7(gpu, port, report, N):
    _check(gpu, port, N, N, po.WIN_BH7, seed=N % 97, report=report)


@pytest.mark.parametrize("wtype", range(7))
def test_windows_64k(gpu, port, wtype):
    _check(gpu, port, 65536, 65536, wtype, seed=3 + wtype)


@pytest.mark.parametrize("N,nz", [(1024, 1000), (4096, 3), (65536, 12000), (131072, 131071), (1048576, 1000000)])
def test_zero_padded(gpu, port, N, nz):
    _check(gpu, port, N, nz, po.WIN_HANN, seed=11)


def test_unity_gain_tone_on_bin(gpu, port):
    # unit tone on bin k -> 0.000 dB at index N/2 + k
    N, k = 65536, 1234
    n = np.arange(N)
    x = np.exp(2j * np.pi * k * n / N).astype(np.complex64)
    x += (1e-6 * (np.cos(0.37 * n) + 1j * np.sin(0.11 * n))).astype(np.complex64)
    row = gpu.spectrum(N, x, port.window(po.WIN_BH7, N))
    assert int(np.argmax(row)) == N // 2 + k
    assert abs(float(row[N // 2 + k])) < 0.002


def test_int_formats_match_converted_input(gpu, port):
    N = 131072
    raw = synth.quantise(synth.baseband(N, 3.2e6, 2), po.FMT_U8_RTL)
    w = port.window(po.WIN_HANN, N)
    row_a = gpu.spectrum(N, raw, w, fmt=po.FMT_U8_RTL)
    row_b = gpu.spectrum(N, port.convert(po.FMT_U8_RTL, raw), w)
    assert np.array_equal(row_a.view(np.uint32), row_b.view(np.uint32))


def test_frontend_framing_saturated(gpu, port):
    """Frames that span block boundaries equal the one-shot transform of the same samples
    (Reshaper keep/skip, reshaper.h:102-129, in saturated mode fftRate = sr/N)."""
    sr, N, blk = 2.4e6, 8192, 12000
    x = _frame(blk * 5, 5)
    with gpu.Frontend(sr, fft_size=N, fft_rate=sr / N, fft_window=po.WIN_BH4, max_block=blk) as fe:
        rows = []
        for b in range(5):
            fe.process(po.FMT_CF32, x[b * blk:(b + 1) * blk])
            rows.append(fe.fft_rows())
        rows = np.concatenate(rows)
    assert rows.shape == (blk * 5 // N, N)
    w = port.window(po.WIN_BH4, N)
    for f in range(rows.shape[0]):
        one = gpu.spectrum(N, x[f * N:(f + 1) * N], w)
        assert np.array_equal(one.view(np.uint32), rows[f].view(np.uint32))


def test_frontend_framing_keep_skip(gpu, port):
    sr, N, blk, rate = 2.4e6, 4096, 12000, 100.0  # interval 24000, nz 4096, skip 19904
    skip, nz = port.reshape_params(sr, N, rate)
    assert (skip, nz) == (19904, 4096)
    x = _frame(blk * 6, 6)
    with gpu.Frontend(sr, fft_size=N, fft_rate=rate, fft_window=po.WIN_NUTTALL, max_block=blk) as fe:
        rows = []
        for b in range(6):
            fe.process(po.FMT_CF32, x[b * blk:(b + 1) * blk])
            rows.append(fe.fft_rows())
        rows = np.concatenate(rows)
    assert rows.shape[0] == 3
    w = port.window(po.WIN_NUTTALL, nz)
    for f in range(3):
        s = f * (nz + skip)
        one = gpu.spectrum(N, x[s:s + nz], w)
        assert np.array_equal(one.view(np.uint32), rows[f].view(np.uint32))


def test_bad_sizes(gpu):
    with pytest.raises(gpu.SdrppCudaError):
        gpu.spectrum(1000, np.zeros(1000, np.complex64), np.ones(1000, np.float32))


def test_low_noise_floor_scatter(gpu, port):
    """-70 dBFS noise puts the per-bin floor near the fp32 round-off of ANY FFT: there the
    GPU row is compared with the reference's own fp32 scatter instead of the 0.01 dB gate."""
    N = 65536
    x = _frame(N, 21, noise=-70.0)
    w = port.window(po.WIN_BH7, N)
    row32, X64, row64 = port.spectrum(N, x, w)
    row, X = gpu.spectrum(N, x, w, want_X=True)
    assert po.rel_rms(X, X64) <= TOL_X
    strong = row64 >= row64.max() - 80.0
    assert np.abs(row.astype(np.float64) - row64)[strong].max() <= TOL_DB
    g = np.abs(row.astype(np.float64) - row64); r = np.abs(row32.astype(np.float64) - row64)
    for q in (99.0, 99.9):
        assert np.percentile(g, q) <= max(SCATTER_X * np.percentile(r, q), TOL_DB), (q, np.percentile(g, q), np.percentile(r, q))


ZOOM_CASES = [(65536, 1024, 0.0, 2.4e6, 2.4e6), (65536, 1600, 3e5, 4e5, 2.4e6), (1048576, 1917, -2e7, 3.3e7, 122.88e6),
              (1024, 2000, 0.0, 2.4e6, 2.4e6), (8192, 777, 1.1e6, 2.4e5, 2.4e6), (8192, 500, -1.19e6, 1e5, 2.4e6)]


@pytest.mark.parametrize("N,out,vo,vb,wb", ZOOM_CASES)
def test_waterfall_zoom(gpu, port, N, out, vo, vb, wb):
    """fft_scaler::doZoom: bin boundaries and pixel values bit-exact."""
    rng = np.random.default_rng(N + out)
    row = (rng.standard_normal(N) * 10.0 - 80.0).astype(np.float32)
    ref, ridx = port.fft_zoom(vo, vb, wb, row, out)
    got, gidx = gpu.fft_zoom(row, vo, vb, wb, out)
    assert np.array_equal(gidx, ridx)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_frontend_zoomed_rows(gpu, port):
    sr, N, blk, W = 2.4e6, 8192, 12000, 1000
    x = _frame(blk * 3, 9)
    with gpu.Frontend(sr, fft_size=N, fft_rate=sr / N, fft_window=po.WIN_BH4, max_block=blk) as fe:
        fe.set_fft_zoom(2e5, 1.2e6, sr, W, keep_raw=True)
        raw, zoomed = [], []
        for b in range(3):
            fe.process(po.FMT_CF32, x[b * blk:(b + 1) * blk])
            raw.append(fe.fft_rows()); zoomed.append(fe.fft_zoomed_rows())
        raw, zoomed = np.concatenate(raw), np.concatenate(zoomed)
    assert raw.shape == (4, N) and zoomed.shape == (4, W)
    for r in range(4):
        ref, _ = port.fft_zoom(2e5, 1.2e6, sr, raw[r], W)
        assert np.array_equal(zoomed[r].view(np.uint32), ref.view(np.uint32))


def test_spectrum_device_batched(gpu, port):
    """sdrpp_cuda_spectrum_device: frames already on the GPU, several per call, equal to the one-shot rows."""
    import torch
    N, F = 65536, 5
    x = _frame(N * F, 41)
    xd = torch.from_numpy(x.view(np.float32).reshape(-1, 2).copy()).cuda()
    rows = torch.empty((F, N), dtype=torch.float32, device="cuda")
    w = port.window(po.WIN_BH7, N)
    gpu.spectrum_device(N, N, F, N, xd.data_ptr(), w, rows.data_ptr())
    torch.cuda.synchronize()
    got = rows.cpu().numpy()
    for f in range(F):
        one = gpu.spectrum(N, x[f * N:(f + 1) * N], w)
        assert np.array_equal(one.view(np.uint32), got[f].view(np.uint32))
