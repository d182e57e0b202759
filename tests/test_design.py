"""Host-side design maths of the product library (sdrpp_cuda_design_*) against the oracle:
window tables, windowed-sinc taps, resampler plans, decimation plans, reshape arithmetic.
These are double->float computations that must match bit for bit."""
import numpy as np
import pytest

from oracle import pyoracle as po

RATES = [(2.4e6, 250e3), (2.4e6, 240e3), (3.2e6, 48e3), (20e6, 250e3), (20e6, 240e3), (15.36e6, 48e3), (15.36e6, 24e3),
         (122.88e6, 48e3), (122.88e6, 24e3), (48e3, 48e3), (48e3, 96e3), (1e6, 300e3), (2.048e6, 1.024e6), (8e6, 8e6)]


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("wtype", range(7))
@pytest.mark.parametrize("n", [1, 2, 8, 1000, 1001, 65536])
def test_window(cuda_lib, port, wtype, n):
    assert np.array_equal(_bits(cuda_lib.design_window(wtype, n)), _bits(port.window(wtype, n)))
    assert np.array_equal(_bits(cuda_lib.design_window(wtype, n, centered=False)), _bits(port.window(wtype, n, centered=False)))


def test_window_known_answer(cuda_lib):
    w = cuda_lib.design_window(po.WIN_BH7, 8)
    ref = np.array([-2.72571e-08, 0.000469102, -0.0293885, 0.239559, -0.461167, 0.239559, -0.0293885, 0.000469102])
    assert np.allclose(w, ref, rtol=2e-5, atol=1e-12)


@pytest.mark.parametrize("spec", [(100e3, 10e3, 250e3), (100e3, 10e3, 240e3), (6250, 625, 48000), (1350, 135, 48000), (6000, 600, 24000), (24000, 2400, 1228800)])
def test_lowpass(cuda_lib, port, spec):
    a, b = cuda_lib.design_lowpass(*spec), port.lowpass_taps(*spec)
    assert len(a) == len(b) and np.array_equal(_bits(a), _bits(b))


@pytest.mark.parametrize("i,o", RATES)
def test_resampler_plan(cuda_lib, port, i, o):
    a, ta = cuda_lib.design_resampler(i, o)
    b, tb = port.resampler_plan(i, o)
    assert a == b
    assert np.array_equal(_bits(ta), _bits(tb))


@pytest.mark.parametrize("k", range(1, 14))
def test_decim_plans(cuda_lib, port, k):
    a, b = cuda_lib.design_decim_plan(1 << k), port.decim_plan(1 << k)
    assert len(a) == len(b) > 0
    for (d1, t1), (d2, t2) in zip(a, b):
        assert d1 == d2 and np.array_equal(_bits(t1), _bits(t2))
    assert int(np.prod([d for d, _ in a])) == 1 << k


def test_decim_plan_invalid(cuda_lib):
    assert cuda_lib.design_decim_plan(3) == [] and cuda_lib.design_decim_plan(1 << 14) == []


@pytest.mark.parametrize("sr,n,rate", [(2.4e6, 65536, 15.0), (3.2e6, 131072, 20.0), (20e6, 1048576, 20.0), (122.88e6, 1048576, 122.88e6 / 1048576), (8e6, 1024, 20.0)])
def test_reshape(cuda_lib, port, sr, n, rate):
    assert cuda_lib.design_reshape(sr, n, rate) == port.reshape_params(sr, n, rate)


@pytest.mark.parametrize("sr", [250e3, 240e3, 192e3])
def test_wfm_pilot_bandpass_bit_exact(cuda_lib, ref, sr):
    """dsp::taps::bandPass<complex_t>(18750, 19250, 3000, sr, true), the pilot filter of dsp::demod::BroadcastFM."""
    got = cuda_lib.design_bandpass_complex(18750.0, 19250.0, 3000.0, sr, odd=True)
    pilot, audio = ref.wfm(75e3, sr).taps()
    assert len(got) == len(pilot) and len(got) % 2 == 1
    assert np.array_equal(got.view(np.uint32), pilot.view(np.uint32))
    assert np.array_equal(cuda_lib.design_lowpass(15000.0, 4000.0, sr), audio)
