"""TEST INFRASTRUCTURE ONLY (oracle). ctypes front ends for the two CPU checkers.

* ``Port``  -> oracle/liboracle_port.so   (our plain-C restatement, oracle/port/oracle.c)
* ``Ref``   -> oracle/_ref/libsdrpp_ref*.so (the reference's own dsp/ headers, oracle/ref_api.cpp)

Both expose the same Python surface so tests can run one against the other and the CUDA path
against either. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm may
import this module; the product package (sdrpp_b200/) never does.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PLANS_BLOB = os.path.join(ROOT, "sdrpp_b200", "data", "decim_plans.bin")

cf32 = np.complex64
_vp, _i, _d = C.c_void_p, C.c_int, C.c_double
_fp = C.POINTER(C.c_float)

# input sample formats (same numbering as include/sdrpp_cuda.h)
FMT_CF32, FMT_U8_RTL, FMT_U8_TCP, FMT_I8, FMT_I16_FILE, FMT_I16_VOLK, FMT_I24_FILE, FMT_I32_FILE, FMT_F64 = range(9)
# window types, dsp/window/window.h:28-36
WIN_RECT, WIN_HAMMING, WIN_HANN, WIN_BLACKMAN, WIN_NUTTALL, WIN_BH4, WIN_BH7 = range(7)
# demod front ends
DEMOD_NONE, DEMOD_QUAD, DEMOD_AM, DEMOD_USB, DEMOD_LSB, DEMOD_DSB = range(6)


def _ptr(a):
    return a.ctypes.data_as(_vp)


def _c64(a):
    return np.ascontiguousarray(a, dtype=cf32)


class _Obj:
    """A block object living inside an oracle library."""

    def __init__(self, lib, prefix, handle, out_dtype=cf32, in_place_ok=True):
        self._lib, self._p, self.h, self._odt = lib, prefix, handle, out_dtype
        if not handle:
            raise RuntimeError(f"{prefix}_create failed")

    def process(self, x, out_cap=None):
        x = _c64(x)
        n = len(x)
        # room for interpolating resamplers (outSR > inSR): the reference writes count*interp/decim samples
        out = np.zeros(max(out_cap or 0, 4 * n, 1) + 64, dtype=self._odt)
        fn = getattr(self._lib, f"{self._p}_process")
        fn.restype = _i
        fn.argtypes = [_vp, _i, _vp, _vp]
        m = fn(self.h, n, _ptr(x), _ptr(out))
        return out[:m].copy()

    def call_void(self, name, *args, argtypes=()):
        fn = getattr(self._lib, f"{self._p}_{name}")
        fn.restype = None
        fn.argtypes = [_vp, *argtypes]
        fn(self.h, *args)

    def reset(self):
        self.call_void("reset")

    def __del__(self):
        try:
            fn = getattr(self._lib, f"{self._p}_destroy")
            fn.restype = None
            fn.argtypes = [_vp]
            if self.h:
                fn(self.h)
                self.h = None
        except Exception:
            pass


class _Base:
    prefix = ""

    def _f(self, name, restype, *argtypes):
        fn = getattr(self.lib, f"{self.prefix}_{name}")
        fn.restype = restype
        fn.argtypes = list(argtypes)
        return fn

    # ---- design maths -------------------------------------------------------------------
    def window(self, wtype, size, centered=True):
        buf = np.zeros(size + 2, dtype=np.float32)
        name = "create_window" if self.prefix == "ref" else "window"
        self._f(name, None if self.prefix == "ref" else _i, _i, _vp, _i, _i)(wtype, _ptr(buf), size, int(centered))
        return buf[:size].copy()

    def lowpass_taps(self, cutoff, trans, sr):
        n = self._f("lowpass_taps", _i, _d, _d, _d, _vp, _i)(cutoff, trans, sr, None, 0)
        out = np.zeros(n, dtype=np.float32)
        self._f("lowpass_taps", _i, _d, _d, _d, _vp, _i)(cutoff, trans, sr, _ptr(out), n)
        return out

    def decim_plan(self, ratio):
        dec, cnt, tp = (_i * 4)(), (_i * 4)(), (_fp * 4)()
        n = self._f("decim_plan", _i, _i, _vp, _vp, _vp)(ratio, dec, cnt, tp)
        return [(dec[i], np.ctypeslib.as_array(tp[i], shape=(cnt[i],)).copy()) for i in range(n)]

    # ---- block objects --------------------------------------------------------------------
    def fir(self, taps):
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        if self.prefix == "ref":
            h = self._f("fir_create", _vp, _vp, _i)(_ptr(taps), len(taps))
        else:
            h = self._f("fir_create", _vp, _vp, _i, _i)(_ptr(taps), len(taps), 1)
        return _Obj(self.lib, f"{self.prefix}_fir", h)

    def decfir(self, taps, decim):
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        if self.prefix == "ref":
            h = self._f("decfir_create", _vp, _vp, _i, _i)(_ptr(taps), len(taps), decim)
            o = _Obj(self.lib, "ref_decfir", h)
            o.offset = lambda: self._f("decfir_offset", _i, _vp)(o.h)
        else:
            h = self._f("fir_create", _vp, _vp, _i, _i)(_ptr(taps), len(taps), decim)
            o = _Obj(self.lib, "orc_fir", h)
            o.offset = lambda: self._f("fir_offset", _i, _vp)(o.h)
        return o

    def powerdecim(self, ratio):
        o = _Obj(self.lib, f"{self.prefix}_powerdecim", self._f("powerdecim_create", _vp, _i)(ratio))

        def offsets():
            arr = (_i * 4)()
            n = self._f("powerdecim_offsets", _i, _vp, _vp)(o.h, arr)
            return [arr[i] for i in range(n)]
        o.offsets = offsets
        return o

    def polyphase(self, interp, decim, taps):
        taps = np.ascontiguousarray(taps, dtype=np.float32)
        o = _Obj(self.lib, f"{self.prefix}_polyphase",
                 self._f("polyphase_create", _vp, _i, _i, _vp, _i)(interp, decim, _ptr(taps), len(taps)))

        def state():
            ph, off = _i(), _i()
            self._f("polyphase_state", None, _vp, _vp, _vp)(o.h, C.byref(ph), C.byref(off))
            return ph.value, off.value
        o.state = state
        return o

    def resampler(self, in_sr, out_sr):
        return _Obj(self.lib, f"{self.prefix}_resampler", self._f("resampler_create", _vp, _d, _d)(in_sr, out_sr))

    def xlator(self, offset_hz, sr, ideal=False):
        """FrequencyXlator; ideal=True: the ideal-NCO flavour (closed-form phase from the fp32-quantised increment)."""
        if ideal and self.prefix == "ref":
            return _Obj(self.lib, "ref_xlatorideal", self._f("xlator_create_ideal", _vp, _d, _d)(offset_hz, sr))
        o = _Obj(self.lib, f"{self.prefix}_xlator", self._f("xlator_create_ideal" if ideal else "xlator_create", _vp, _d, _d)(offset_hz, sr))

        def state():
            ph, dl = (C.c_float * 2)(), (C.c_float * 2)()
            self._f("xlator_state", None, _vp, _vp, _vp)(o.h, ph, dl)
            return (ph[0], ph[1]), (dl[0], dl[1])
        o.state = state
        o.set_offset = lambda off, sr_: self._f("xlator_set_offset", None, _vp, _d, _d)(o.h, off, sr_)
        return o

    def rxvfo(self, in_sr, out_sr, bw, offset, ideal_nco=False):
        """RxVFO; ideal_nco=True: same object with the xlator's fp32 phase recurrence replaced by the closed form."""
        o = _Obj(self.lib, f"{self.prefix}_rxvfo",
                 self._f("rxvfo_create_ideal" if ideal_nco else "rxvfo_create", _vp, _d, _d, _d, _d)(in_sr, out_sr, bw, offset))
        o.set_offset = lambda off: self._f("rxvfo_set_offset", None, _vp, _d)(o.h, off)

        def info():
            arr = (_i * 8)()
            if self.prefix == "ref":
                self._f("rxvfo_info", None, _vp, _vp, _vp, _i, _vp, _i)(o.h, arr, None, 0, None, 0)
            else:
                self._f("rxvfo_info", _i, _vp, _vp)(o.h, arr)
            return dict(mode=arr[0], predec=arr[1], interp=arr[2], decim=arr[3], rtaps=arr[4], tpp=arr[5], ftaps=arr[6])
        o.info = info
        return o

    def noise_blanker(self, rate, level):
        return _Obj(self.lib, f"{self.prefix}_nb", self._f("nb_create", _vp, _d, _d)(rate, level))

    def squelch(self, level):
        return _Obj(self.lib, f"{self.prefix}_squelch", self._f("squelch_create", _vp, _d)(level))

    def fm_if(self, bins):
        return _Obj(self.lib, f"{self.prefix}_fmif", self._f("fmif_create", _vp, _i)(bins))

    def dcblock(self, rate):
        return _Obj(self.lib, f"{self.prefix}_dcblock", self._f("dcblock_create", _vp, _d)(rate))

    def quadrature(self, deviation, sr):
        return _Obj(self.lib, f"{self.prefix}_quadrature", self._f("quadrature_create", _vp, _d, _d)(deviation, sr), out_dtype=np.float32)

    def ssb(self, mode, bw, sr, ideal_nco=False):
        return _Obj(self.lib, f"{self.prefix}_ssb", self._f("ssb_create_ideal" if ideal_nco else "ssb_create", _vp, _i, _d, _d)(mode, bw, sr), out_dtype=np.float32)

    def am_magnitude(self, x):
        x = _c64(x)
        out = np.zeros(len(x), dtype=np.float32)
        self._f("am_magnitude", _i, _i, _vp, _vp)(len(x), _ptr(x), _ptr(out))
        return out

    def conjugate(self, x):
        x = _c64(x)
        out = np.zeros_like(x)
        self._f("conjugate", _i, _i, _vp, _vp)(len(x), _ptr(x), _ptr(out))
        return out

    # ---- complete demodulators -----------------------------------------------------
    def _post_obj(self, h, name):
        lib, pre = self.lib, self.prefix
        proc = f"{pre}_{name}" if pre == "ref" else "orc_post"
        o = _Obj(lib, proc, h, out_dtype=np.float32)
        return o

    def fm_full(self, sr, bw, low_pass=True):
        return self._post_obj(self._f("fm_create", _vp, _d, _d, _i)(sr, bw, int(low_pass)), "fm")

    def am_full(self, agc_mode, bw, attack, decay, dc_rate, sr, agc_gain=0.0):
        return self._post_obj(self._f("am_create", _vp, _i, _d, _d, _d, _d, _d, C.c_float)(agc_mode, bw, attack, decay, dc_rate, sr, agc_gain), "am")

    def ssb_full(self, mode, bw, sr, agc_enabled, attack, decay):
        return self._post_obj(self._f("ssbfull_create", _vp, _i, _d, _d, _i, _d, _d)(mode, bw, sr, int(agc_enabled), attack, decay), "ssbfull")

    def wfm(self, deviation, sr, stereo=True, low_pass=True, rds=False, ideal_nco=False):
        """dsp::demod::BroadcastFM: process(iq) -> interleaved (l, r) floats, 2 per input sample. rds=True: the decoder's RDS side
        output is on and process returns (lr, rds) -- rds = complex samples at 5 kS/s produced by this call. ideal_nco (port only):
        the -57 kHz translation with the closed-form phase instead of the fp32 rotator."""
        lib, pre = self.lib, self.prefix
        if rds and pre == "orc":
            h = self._f("wfm_create_rds", _vp, _d, _d, _i, _i, _i)(deviation, sr, int(stereo), int(low_pass), 2 if ideal_nco else 1)
        elif rds:
            if ideal_nco:
                raise ValueError("the reference build has no ideal-NCO BroadcastFM (its xlator is a private member)")
            h = self._f("wfm_create_rds", _vp, _d, _d, _i, _i)(deviation, sr, int(stereo), int(low_pass))
        else:
            h = self._f("wfm_create", _vp, _d, _d, _i, _i)(deviation, sr, int(stereo), int(low_pass))
        o = _Obj(lib, f"{pre}_wfm", h, out_dtype=np.float32)
        base = self

        def process(x, out_cap=None):
            x = _c64(x)
            out = np.zeros(2 * len(x) + 16, dtype=np.float32)
            if rds:
                r = np.zeros(len(x) + 16, dtype=cf32)
                nr = _i(0)
                fn = getattr(lib, f"{pre}_wfm_process_rds"); fn.restype = _i; fn.argtypes = [_vp, _i, _vp, _vp, _vp, _vp]
                n = fn(o.h, len(x), _ptr(x), _ptr(out), _ptr(r), C.byref(nr))
                return out[:2 * n].reshape(n, 2).copy(), r[:nr.value].copy()
            fn = getattr(lib, f"{pre}_wfm_process"); fn.restype = _i; fn.argtypes = [_vp, _i, _vp, _vp]
            n = fn(o.h, len(x), _ptr(x), _ptr(out))
            return out[:2 * n].reshape(n, 2).copy()
        o.process = process

        def taps():
            n = (_i * 2)()
            base._f("wfm_taps", None, _vp, _vp, _vp, _i, _vp, _i)(o.h, n, None, 0, None, 0)
            p, a = np.zeros(2 * n[0], np.float32), np.zeros(n[1], np.float32)
            base._f("wfm_taps", None, _vp, _vp, _vp, _i, _vp, _i)(o.h, n, _ptr(p), n[0], _ptr(a), n[1])
            return p.view(np.complex64), a
        o.taps = taps
        return o

    def demod(self, kind, bw, sr, ideal_nco=False):
        """Demod front end object for a VFO output stream (kind = DEMOD_*), or None."""
        if kind == DEMOD_QUAD:
            return self.quadrature(bw / 2.0, sr)
        if kind in (DEMOD_USB, DEMOD_LSB, DEMOD_DSB):
            return self.ssb({DEMOD_USB: 0, DEMOD_LSB: 1, DEMOD_DSB: 2}[kind], bw, sr, ideal_nco=ideal_nco)
        if kind == DEMOD_AM:
            class _AM:
                def process(s, x, out_cap=None):
                    return self.am_magnitude(x)
            return _AM()
        return None

    # ---- spectrum -----------------------------------------------------------------------
    def spectrum(self, N, frame, window, want64=True):
        frame = _c64(frame)
        window = np.ascontiguousarray(window, dtype=np.float32)
        nz = len(window)
        assert len(frame) >= nz
        row32 = np.zeros(N, dtype=np.float32)
        X64 = np.zeros(N, dtype=np.complex128) if want64 else None
        row64 = np.zeros(N, dtype=np.float64) if want64 else None
        rc = self._f("spectrum", _i, _i, _i, _vp, _vp, _vp, _vp, _vp)(
            N, nz, _ptr(frame), _ptr(window), _ptr(row32), _ptr(X64) if want64 else None, _ptr(row64) if want64 else None)
        if rc != 0:
            raise ValueError("spectrum: bad size")
        return row32, X64, row64

    # SDR++ server wire packets (dsp/compression/sample_stream_{de,}compressor.h)
    def pcm_decompress(self, packet):
        packet = np.ascontiguousarray(packet, dtype=np.uint8)
        out = np.zeros(max(1, (len(packet) - 8) // 2), dtype=np.complex64)
        n = self._f("pcm_decompress", _i, _i, _vp, _vp)(len(packet), _ptr(packet), _ptr(out))
        return out[:n].copy()

    def pcm_compress(self, pcm_type, x):
        x = np.ascontiguousarray(x, dtype=np.complex64)
        packet = np.zeros(8 + 8 * len(x), dtype=np.uint8)
        n = self._f("pcm_compress", _i, _i, _i, _vp, _vp)(len(x), int(pcm_type), _ptr(x), _ptr(packet))
        return packet[:n].copy()


def _zoom(fn, view_offset, view_bw, whole_bw, row, out_size, with_idx):
    row = np.ascontiguousarray(row, dtype=np.float32)
    n = len(row)
    data = np.concatenate([row, np.zeros(1, np.float32)])   # doZoom may touch data[fftSize]
    out = np.zeros(out_size, dtype=np.float32)
    if with_idx:
        idx = np.zeros(out_size + 1, dtype=np.int32)
        fn(view_offset, view_bw, whole_bw, n, out_size, _ptr(data), _ptr(out), _ptr(idx))
        return out, idx
    fn(view_offset, view_bw, whole_bw, n, out_size, _ptr(data), _ptr(out))
    return out


class Port(_Base):
    """Plain-C restatement (oracle/port/oracle.c)."""
    prefix = "orc"

    def __init__(self, flavour=""):
        """flavour '' = fp32 (the parity oracle); 'f64' = fp64-accumulated dot products (scatter adjudication only)."""
        path = os.path.join(HERE, "liboracle_port.so" if not flavour else f"liboracle_port_{flavour}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle port` (or __graft_entry__.build())")
        self.lib = C.CDLL(path)
        rc = self._f("load_plans", _i, C.c_char_p)(PLANS_BLOB.encode())
        if rc != 0:
            raise RuntimeError(f"orc_load_plans({PLANS_BLOB}) -> {rc}")

    def convert(self, fmt, raw):
        raw = np.ascontiguousarray(raw)
        n = raw.size * 2 if fmt == FMT_CF32 else (raw.size // 3 if fmt == FMT_I24_FILE else raw.size)
        out = np.zeros(n, dtype=np.float32)
        rc = self._f("convert", _i, _i, _vp, _i, _vp)(fmt, _ptr(raw), n, _ptr(out))
        if rc != 0:
            raise ValueError("bad format")
        return out.view(cf32)

    def fft_zoom(self, view_offset, view_bw, whole_bw, row, out_size):
        """fft_scaler::doZoom; returns (pixels, bin boundaries)."""
        return _zoom(self._f("fft_zoom", None, _d, _d, _d, _i, _i, _vp, _vp, _vp), view_offset, view_bw, whole_bw, row, out_size, True)

    def vfo_signal_info(self, row, center_offset, bandwidth, whole_bw):
        """WaterFall::calculateVFOSignalInfo (waterfall.cpp:563-603): (strength, snr) of one raw row."""
        row = np.ascontiguousarray(row, dtype=np.float32)
        st, sn = C.c_float(0), C.c_float(0)
        self._f("vfo_signal_info", _i, _vp, _i, _d, _d, _d, _vp, _vp)(_ptr(row), len(row), center_offset, bandwidth, whole_bw, C.byref(st), C.byref(sn))
        return st.value, sn.value

    def fft_display(self, rows, smoothing, alpha, smooth_buf, hold, hold_speed, hold_buf):
        """WaterFall::pushFFT smoothing + peak hold on zoomed rows (in place on copies); returns (rows, smooth_buf, hold_buf)."""
        rows = np.array(rows, dtype=np.float32, copy=True)
        sb = np.array(smooth_buf, dtype=np.float32, copy=True)
        hb = np.array(hold_buf, dtype=np.float32, copy=True)
        self._f("fft_display", None, _i, _i, _vp, _i, C.c_float, _vp, _i, C.c_float, _vp)(rows.shape[1], rows.shape[0], _ptr(rows), int(smoothing), alpha, _ptr(sb),
                                                                                           int(hold), hold_speed, _ptr(hb))
        return rows, sb, hb

    def reshape_params(self, sr, size, rate):
        skip, nz = _i(), _i()
        self._f("reshape_params", None, _d, _i, _d, _vp, _vp)(sr, size, rate, C.byref(skip), C.byref(nz))
        return skip.value, nz.value

    def resampler_plan(self, in_sr, out_sr):
        info = (_i * 6)()
        self._f("resampler_plan", _i, _d, _d, _vp, _vp, _i)(in_sr, out_sr, info, None, 0)
        taps = np.zeros(info[4], dtype=np.float32)
        if info[4]:
            self._f("resampler_plan", _i, _d, _d, _vp, _vp, _i)(in_sr, out_sr, info, _ptr(taps), info[4])
        return dict(mode=info[0], predec=info[1], interp=info[2], decim=info[3], ntaps=info[4], tpp=info[5]), taps


class Ref(_Base):
    """The reference's own dsp/ headers (oracle/_ref). flavour: '' (fp32), 'f64', 'fast'."""
    prefix = "ref"

    def __init__(self, flavour=""):
        name = "libsdrpp_ref.so" if not flavour else f"libsdrpp_ref_{flavour}.so"
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(path)
        self.flavour = flavour

    def fft_zoom(self, view_offset, view_bw, whole_bw, row, out_size):
        return _zoom(self._f("fft_zoom", None, _d, _d, _d, _i, _i, _vp, _vp), view_offset, view_bw, whole_bw, row, out_size, False)

    def resampler_info(self, obj):
        info = (_i * 6)()
        self._f("resampler_info", None, _vp, _vp, _vp, _i)(obj.h, info, None, 0)
        taps = np.zeros(info[4], dtype=np.float32)
        if info[4]:
            self._f("resampler_info", None, _vp, _vp, _vp, _i)(obj.h, info, _ptr(taps), info[4])
        return dict(mode=info[0], predec=info[1], interp=info[2], decim=info[3], ntaps=info[4], tpp=info[5]), taps

    def bench_channelizer(self, in_sr, vfos, block, nblocks, nthreads, fft=None):
        """vfos: list of (outSR, bw, offset, demod). fft: None or (N, window, frames); the block
        buffer must then hold at least len(window) samples. Returns elapsed seconds."""
        V = len(vfos)
        arr = lambda k, t: (t * V)(*[v[k] for v in vfos])
        block = _c64(block)
        count = len(block)
        fn = self._f("bench_channelizer", _d, _d, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _vp, _i)
        if fft is not None:
            N, win, frames = fft
            win = np.ascontiguousarray(win, dtype=np.float32)
            assert count >= len(win), "block buffer shorter than the FFT frame"
            return fn(in_sr, V, arr(0, _d), arr(1, _d), arr(2, _d), arr(3, _i), count, nblocks, nthreads, _ptr(block),
                      N, len(win), _ptr(win), frames)
        return fn(in_sr, V, arr(0, _d), arr(1, _d), arr(2, _d), arr(3, _i), count, nblocks, nthreads, _ptr(block), 0, 0, None, 0)


def have_ref(flavour=""):
    name = "libsdrpp_ref.so" if not flavour else f"libsdrpp_ref_{flavour}.so"
    return os.path.exists(os.path.join(HERE, "_ref", name))


def fnv1a_words(a):
    """Word-wise FNV-1a over the u32 bit patterns of an array (SURVEY 8c known answers)."""
    h = 2166136261
    for w in np.ascontiguousarray(a).view(np.uint32).ravel().tolist():
        h = ((h ^ w) * 16777619) & 0xFFFFFFFF
    return h


def rel_rms(a, b):
    a = np.asarray(a); b = np.asarray(b)
    den = np.sqrt(np.mean(np.abs(b.astype(np.complex128)) ** 2))
    return float(np.sqrt(np.mean(np.abs(a.astype(np.complex128) - b.astype(np.complex128)) ** 2)) / (den if den > 0 else 1.0))


def aligned_rel_rms(a, b):
    """Phase-aligned residual: fit the complex scalar c minimising |a - c*b|."""
    a = np.asarray(a, dtype=np.complex128); b = np.asarray(b, dtype=np.complex128)
    c = np.vdot(b, a) / max(np.vdot(b, b).real, 1e-300)
    return rel_rms(a, c * b), c
