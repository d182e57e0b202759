"""GPU parity of the waterfall's consumers of a spectrum row: WaterFall::calculateVFOSignalInfo
(level / SNR, gui/widgets/waterfall.cpp:563-603 -- what the scanner module reads) and the per-line display state of
WaterFall::pushFFT on the zoomed row (FFT smoothing and peak hold, waterfall.cpp:918-956), against the oracle's
restatement (the widget itself needs ImGui / OpenGL and cannot be compiled here)."""
import numpy as np
import pytest

from oracle import pyoracle as po
from sdrpp_b200 import synth

pytestmark = pytest.mark.gpu


def test_signal_info_one_shot(gpu, port, report):
    N, sr = 65536, 2.4e6
    x = synth.baseband(N, sr, 41, carriers=[(300e3, "fm"), (-500e3, "am")], noise_dbfs=-50.0).astype(np.complex64)
    row = gpu.spectrum(N, x, port.window(po.WIN_BH7, N))
    # VFOs on a carrier, on noise, narrow, wide, and touching both band edges
    co = np.array([300e3, -500e3, 0.0, 123456.7, -1.19e6, 1.19e6, 1.0e6], dtype=np.float64)
    bw = np.array([200e3, 12.5e3, 2700.0, 50e3, 100e3, 100e3, 1.0e6], dtype=np.float64)
    st, sn = gpu.signal_info(row, co, bw, sr)
    worst = 0.0
    for i in range(len(co)):
        rs, rn = port.vfo_signal_info(row, co[i], bw[i], sr)
        assert st[i] == np.float32(rs), (i, st[i], rs)                      # a maximum: bit-exact
        assert abs(float(sn[i]) - rn) <= 2e-5, (i, sn[i], rn)                # mean of up to N/2 bins in double, other order
        worst = max(worst, abs(float(sn[i]) - rn))
    report("8f-2 calculateVFOSignalInfo one-shot", strength_bit_exact=True, worst_snr_abs_db=worst)


def test_frontend_signal_info_smoothing_and_hold(gpu, port, report):
    sr, blk, N, W = 2.4e6, 12000, 8192, 1000
    nblocks = 30
    offs = [300e3, -500e3]
    x = synth.baseband(blk * nblocks, sr, 42, carriers=[(offs[0], "fm"), (offs[1], "am")], noise_dbfs=-50.0).astype(np.complex64)
    view = (100e3, 1.2e6, sr)
    alpha, hold_speed, snr_alpha = 0.3, 0.25, 0.4
    with gpu.Frontend(sr, fft_size=N, fft_rate=sr / N, fft_window=po.WIN_BH7, max_block=blk) as fe:
        ids = [fe.add_vfo(240e3, 200e3, offs[0]), fe.add_vfo(24e3, 12e3, offs[1], po.DEMOD_AM)]
        fe.set_fft_zoom(*view, W, keep_raw=True)
        fe.set_fft_display(smoothing=True, smoothing_speed=alpha, hold=True, hold_speed=hold_speed)
        for v in ids:
            fe.vfo_set_signal_info(v, True)
        fe.set_snr_smoothing(True, snr_alpha)
        sb = np.full(W, -1000.0, np.float32)      # waterfall.cpp:774-788: buffers start at -1000 dB
        hb = np.full(W, -1000.0, np.float32)
        snr_state = [0.0, 0.0]
        levels = [[], []]
        rows_seen = 0
        for b in range(nblocks):
            if b == 15:
                fe.vfo_set_offset(ids[1], -400e3)   # the read-out follows a retune
                offs[1] = -400e3
            fe.process(po.FMT_CF32, x[b * blk:(b + 1) * blk])
            raw = fe.fft_rows()
            zr = fe.fft_zoomed_rows()
            assert len(raw) == len(zr)
            if len(raw) == 0:
                continue
            rows_seen += len(raw)
            zoomed = np.stack([port.fft_zoom(*view, r, W)[0] for r in raw])
            want, sb, hb = port.fft_display(zoomed, True, alpha, sb, True, hold_speed, hb)
            assert np.array_equal(zr.view(np.uint32), want.view(np.uint32)), f"block {b}: smoothed rows"
            assert np.array_equal(fe.fft_hold_row().view(np.uint32), hb.view(np.uint32)), f"block {b}: hold row"
            for k, vid in enumerate(ids):
                st, sn, lmax = fe.vfo_signal_info(vid)
                assert len(st) == len(raw)
                bwk = 200e3 if k == 0 else 12e3
                for r in range(len(raw)):
                    rs, rn = port.vfo_signal_info(raw[r], offs[k], bwk, sr)
                    snr_state[k] = np.float32((np.float32(1.0) - np.float32(snr_alpha)) * np.float32(snr_state[k])) + np.float32(np.float32(snr_alpha) * np.float32(rn))
                    levels[k] = (levels[k] + [rs])[-10:]
                    assert st[r] == np.float32(rs)
                    assert abs(float(sn[r]) - float(snr_state[k])) <= 5e-5, (b, k, r, sn[r], snr_state[k])
                assert lmax == np.float32(max(levels[k]))
        assert rows_seen == (blk * nblocks) // N
    report("8f-2 pushFFT smoothing + hold, signal info in the front end", rows=rows_seen, smoothed_rows_bit_exact=True, hold_row_bit_exact=True)
