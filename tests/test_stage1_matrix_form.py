"""Host-side model of the tensor-core stage 1 (sdrpp_b200/csrc/channelizer_tc.cu), no GPU needed.

Checks the algebra the kernel relies on, in numpy: (1) NCO + decimating FIR (frequency_xlator.h:43-50 +
decimating_fir.h:45-68) equals the row-matrix form y[m] = e^{j phi(R_m)} * sum_a (X B)[R_m + a][a] with the taps
shifted by the window's position inside a row, for any row origin; (2) the fp16 hi/lo split with a power-of-two
block scale and three products (Xhi*Bhi + Xhi*Blo + Xlo*Bhi, fp32 accumulation) keeps the result within 1e-6 of the
fp64 sum -- the 1e-5 gate of SURVEY 8d leaves an order of magnitude."""
import numpy as np
import pytest


def direct(x, h, D, n0, M, w):
    """y[m] = sum_k h[k] x[n0 + m D + k] e^{j w (n0 + m D + k)}"""
    T = len(h)
    n = n0 + np.arange(M)[:, None] * D + np.arange(T)[None, :]
    return np.sum(h[None, :] * x[n] * np.exp(1j * w * n), axis=1)


def split16(v):
    """fp16 hi/lo of a real array scaled so that its largest component lies in [2^13, 2^14)."""
    m = np.max(np.abs(v))
    e = 14 - int(np.frexp(m)[1]) if m > 0 else 0
    s = v * np.float64(2.0 ** e)
    hi = s.astype(np.float32).astype(np.float16)
    lo = (s.astype(np.float32) - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32), 2.0 ** -e


def matrix_form(x, h, D, n0, M, w, origin, split):
    T = len(h)
    s = (n0 - origin) % D
    A = -(-(T + s) // D)
    R0 = (n0 - origin) // D                      # row in which the window of output 0 starts
    hp = np.zeros(A * D)
    hp[s:s + T] = h
    k = np.arange(A * D)
    Bc = (hp * np.exp(1j * w * k)).reshape(A, D)  # [a][p]
    rows = R0 + np.arange(M + A - 1)
    Xc = x[origin + rows[:, None] * D + np.arange(D)[None, :]]  # [row][p], untranslated samples
    # real form: K = (p, re/im), N = (a, re/im)
    Xr = np.empty((len(rows), 2 * D)); Xr[:, 0::2] = Xc.real; Xr[:, 1::2] = Xc.imag
    Br = np.empty((2 * A, 2 * D))
    Br[0::2, 0::2] = Bc.real; Br[0::2, 1::2] = -Bc.imag   # real output column
    Br[1::2, 0::2] = Bc.imag; Br[1::2, 1::2] = Bc.real    # imaginary output column
    if split:
        V = np.zeros((len(rows), 2 * A), dtype=np.float32)
        bh, bl, bs = split16(Br)
        for g0 in range(0, len(rows), 8):             # block scale per 8 rows, like s1t_split_kernel
            xh, xl, xs = split16(Xr[g0:g0 + 8])
            acc = (xh @ bh.T + xh @ bl.T + xl @ bh.T).astype(np.float32)  # fp32 accumulation (float32 matmul)
            V[g0:g0 + 8] = acc * np.float32(xs * bs)
        V = V.astype(np.float64)
    else:
        V = Xr @ Br.T
    Vc = V[:, 0::2] + 1j * V[:, 1::2]               # [row][a]
    y = np.zeros(M, dtype=np.complex128)
    for a in range(A):
        y += Vc[a:a + M, a]
    return y * np.exp(1j * w * (origin + (R0 + np.arange(M)) * D))


@pytest.mark.parametrize("D,T", [(64, 329), (64, 400), (32, 143), (32, 129), (64, 257)])
@pytest.mark.parametrize("origin", [0, 4, 48])
def test_matrix_form_equals_direct(D, T, origin):
    rng = np.random.default_rng(D * 1000 + T + origin)
    h = np.sinc((np.arange(T) - (T - 1) / 2) / D) * np.hanning(T) / D
    M = 40
    n0 = origin + 17 + 3 * D                          # arbitrary window start
    x = (rng.standard_normal(n0 + M * D + T + 8 * D) + 1j * rng.standard_normal(n0 + M * D + T + 8 * D)) * 0.3
    w = 2 * np.pi * 0.1234567
    ref = direct(x, h, D, n0, M, w)
    exact = matrix_form(x, h, D, n0, M, w, origin, split=False)
    assert np.max(np.abs(exact - ref)) <= 1e-12 * np.max(np.abs(ref)) + 1e-13
    got = matrix_form(x.astype(np.complex64).astype(np.complex128), h.astype(np.float32).astype(np.float64), D, n0, M, w, origin, split=True)
    ref32 = direct(x.astype(np.complex64).astype(np.complex128), h.astype(np.float32).astype(np.float64), D, n0, M, w)
    err = np.sqrt(np.mean(np.abs(got - ref32) ** 2) / np.mean(np.abs(ref32) ** 2))
    assert err <= 1e-6, f"split-fp16 relative RMS {err:.2e}"


@pytest.mark.parametrize("scale", [1e-7, 1.0, 1e4])
def test_block_scale_keeps_precision(scale):
    """The block exponent follows the data: the relative error does not depend on the input amplitude."""
    rng = np.random.default_rng(7)
    D, T, M = 64, 329, 24
    h = np.sinc((np.arange(T) - (T - 1) / 2) / D) * np.hanning(T) / D
    n0 = 5 * D + 9
    x = (rng.standard_normal(n0 + M * D + T + 8 * D) + 1j * rng.standard_normal(n0 + M * D + T + 8 * D)) * scale
    x = x.astype(np.complex64).astype(np.complex128)
    h = h.astype(np.float32).astype(np.float64)
    ref = direct(x, h, D, n0, M, 0.7)
    got = matrix_form(x, h, D, n0, M, 0.7, 8, split=True)
    err = np.sqrt(np.mean(np.abs(got - ref) ** 2) / np.mean(np.abs(ref) ** 2))
    assert err <= 1e-6, f"{err:.2e}"
