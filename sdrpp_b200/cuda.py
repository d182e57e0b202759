"""ctypes binding of include/sdrpp_cuda.h (sdrpp_b200/libsdrpp_cuda.so).

This is the Python face of the C ABI used by the tests and bench.py; the product host side is the
C++ header mirror in include/sdrpp/. There is no CPU fallback: if the library is missing or no CUDA
device is present, calls raise.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDRPP_CUDA_LIB") or os.path.join(HERE, "libsdrpp_cuda.so")  # override: experiment builds

FMT_CF32, FMT_U8_RTL, FMT_U8_TCP, FMT_I8, FMT_I16_FILE, FMT_I16_VOLK, FMT_I24_FILE, FMT_I32_FILE, FMT_F64 = range(9)
WIN_RECT, WIN_HAMMING, WIN_HANN, WIN_BLACKMAN, WIN_NUTTALL, WIN_BH4, WIN_BH7 = range(7)
DEMOD_NONE, DEMOD_QUAD, DEMOD_AM, DEMOD_USB, DEMOD_LSB, DEMOD_DSB = range(6)

FMT_DTYPE = {FMT_CF32: np.complex64, FMT_U8_RTL: np.uint8, FMT_U8_TCP: np.uint8, FMT_I8: np.int8,
             FMT_I16_FILE: np.int16, FMT_I16_VOLK: np.int16, FMT_I24_FILE: np.uint8, FMT_I32_FILE: np.int32, FMT_F64: np.float64}
FMT_BYTES = {FMT_CF32: 8, FMT_U8_RTL: 2, FMT_U8_TCP: 2, FMT_I8: 2, FMT_I16_FILE: 4, FMT_I16_VOLK: 4, FMT_I24_FILE: 6, FMT_I32_FILE: 8, FMT_F64: 16}

# every symbol include/sdrpp_cuda.h declares (checked by tests/test_abi.py)
SYMBOLS = """
sdrpp_cuda_version sdrpp_cuda_last_error sdrpp_cuda_device_count sdrpp_cuda_init sdrpp_cuda_host_alloc
sdrpp_cuda_host_free sdrpp_cuda_design_window sdrpp_cuda_design_lowpass sdrpp_cuda_design_resampler
sdrpp_cuda_design_decim_plan sdrpp_cuda_design_reshape sdrpp_cuda_convert sdrpp_cuda_spectrum
sdrpp_cuda_pcm_decompress sdrpp_cuda_pcm_compress sdrpp_cuda_frontend_submit_pcm
sdrpp_cuda_frontend_create sdrpp_cuda_frontend_destroy sdrpp_cuda_frontend_set_sample_rate
sdrpp_cuda_frontend_set_decimation sdrpp_cuda_frontend_set_dc_blocking sdrpp_cuda_frontend_set_invert_iq
sdrpp_cuda_frontend_set_fft_size sdrpp_cuda_frontend_set_fft_rate sdrpp_cuda_frontend_set_fft_window
sdrpp_cuda_frontend_effective_samplerate sdrpp_cuda_vfo_create sdrpp_cuda_vfo_destroy sdrpp_cuda_vfo_set_offset
sdrpp_cuda_vfo_set_bandwidth sdrpp_cuda_vfo_set_out_samplerate sdrpp_cuda_vfo_reset sdrpp_cuda_vfo_info
sdrpp_cuda_frontend_submit sdrpp_cuda_frontend_submit_device sdrpp_cuda_frontend_wait
sdrpp_cuda_frontend_set_readback sdrpp_cuda_vfo_output sdrpp_cuda_fft_rows sdrpp_cuda_frontend_read_iq
sdrpp_cuda_frontend_launches sdrpp_cuda_frontend_stream sdrpp_cuda_frontend_set_profiling
sdrpp_cuda_frontend_kernel_ms sdrpp_cuda_fft_zoom sdrpp_cuda_frontend_set_fft_zoom sdrpp_cuda_fft_zoomed_rows sdrpp_cuda_spectrum_device
sdrpp_cuda_vfo_set_post sdrpp_cuda_vfo_audio sdrpp_cuda_vfo_set_if_chain sdrpp_cuda_vfo_squelch_state
sdrpp_cuda_frontend_set_stage1_mode sdrpp_cuda_frontend_stage1_tensor_launches
sdrpp_cuda_frontend_set_graphs sdrpp_cuda_frontend_graph_stats
sdrpp_cuda_frontend_wait_input sdrpp_cuda_frontend_pending sdrpp_cuda_frontend_drain sdrpp_cuda_frontend_join_streams
sdrpp_cuda_comm_unique_id sdrpp_cuda_comm_create sdrpp_cuda_comm_destroy sdrpp_cuda_comm_info sdrpp_cuda_frontend_set_comm
sdrpp_cuda_frontend_submit_shared sdrpp_cuda_frontend_set_fft_display sdrpp_cuda_fft_hold_row sdrpp_cuda_vfo_set_signal_info
sdrpp_cuda_frontend_set_snr_smoothing sdrpp_cuda_vfo_signal_info sdrpp_cuda_signal_info sdrpp_cuda_vfo_audio_stereo sdrpp_cuda_vfo_rds sdrpp_cuda_design_bandpass_complex
""".split()

_vp, _i, _d = C.c_void_p, C.c_int, C.c_double


class FrontendCfg(C.Structure):
    _fields_ = [("sample_rate", _d), ("decim_ratio", _i), ("dc_blocking", _i), ("invert_iq", _i),
                ("fft_size", _i), ("fft_rate", _d), ("fft_window", _i), ("max_block", _i),
                ("ring_log2", _i), ("max_fft_rows", _i)]


class PostCfg(C.Structure):
    _fields_ = [("enabled", _i), ("fm_lowpass", _i), ("am_agc_mode", _i), ("ssb_agc", _i),
                ("agc_attack", _d), ("agc_decay", _d), ("dc_block_rate", _d), ("agc_gain", C.c_float), ("wfm", _i), ("wfm_stereo", _i), ("wfm_rds", _i)]


class IfCfg(C.Structure):
    _fields_ = [("nb_enabled", _i), ("nb_rate", _d), ("nb_level", _d), ("squelch_enabled", _i), ("squelch_level", _d),
                ("fmif_bins", _i)]


class SdrppCudaError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SdrppCudaError(f"{LIB_PATH} is missing: run `python -m sdrpp_b200.build` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.sdrpp_cuda_version.restype = C.c_char_p
        L.sdrpp_cuda_last_error.restype = C.c_char_p
        L.sdrpp_cuda_host_alloc.restype = _vp
        L.sdrpp_cuda_host_alloc.argtypes = [C.c_size_t]
        L.sdrpp_cuda_host_free.argtypes = [_vp]
        L.sdrpp_cuda_design_window.argtypes = [_i, _vp, _i, _i]
        L.sdrpp_cuda_design_lowpass.argtypes = [_d, _d, _d, _vp, _i]
        L.sdrpp_cuda_design_resampler.argtypes = [_d, _d, _vp, _vp, _i]
        L.sdrpp_cuda_design_bandpass_complex.argtypes = [_d, _d, _d, _d, _i, _vp, _i]
        L.sdrpp_cuda_design_decim_plan.argtypes = [_i, _vp, _vp, _vp]
        L.sdrpp_cuda_design_reshape.argtypes = [_d, _i, _d, _vp, _vp]
        L.sdrpp_cuda_design_reshape.restype = None
        L.sdrpp_cuda_convert.argtypes = [_i, _vp, _i, _vp]
        L.sdrpp_cuda_spectrum.argtypes = [_i, _i, _i, _vp, _vp, _vp, _vp]
        L.sdrpp_cuda_pcm_decompress.argtypes = [_vp, _i, _vp]
        L.sdrpp_cuda_pcm_compress.argtypes = [_i, _vp, _i, _vp]
        L.sdrpp_cuda_frontend_submit_pcm.argtypes = [_vp, _vp, _i]
        L.sdrpp_cuda_spectrum_device.argtypes = [_i, _i, _i, C.c_longlong, _vp, _vp, _vp, _vp]
        L.sdrpp_cuda_fft_zoom.argtypes = [_i, _vp, _d, _d, _d, _i, _vp, _vp]
        L.sdrpp_cuda_frontend_set_fft_zoom.argtypes = [_vp, _d, _d, _d, _i, _i]
        L.sdrpp_cuda_fft_zoomed_rows.argtypes = [_vp, C.POINTER(_vp)]
        L.sdrpp_cuda_frontend_create.restype = _vp
        L.sdrpp_cuda_frontend_create.argtypes = [C.POINTER(FrontendCfg)]
        L.sdrpp_cuda_frontend_destroy.argtypes = [_vp]
        L.sdrpp_cuda_frontend_set_sample_rate.argtypes = [_vp, _d]
        L.sdrpp_cuda_frontend_set_fft_rate.argtypes = [_vp, _d]
        for n in ("set_decimation", "set_dc_blocking", "set_invert_iq", "set_fft_size", "set_fft_window",
                  "set_readback", "set_profiling"):
            getattr(L, "sdrpp_cuda_frontend_" + n).argtypes = [_vp, _i]
        L.sdrpp_cuda_frontend_effective_samplerate.restype = _d
        L.sdrpp_cuda_frontend_effective_samplerate.argtypes = [_vp]
        L.sdrpp_cuda_vfo_create.argtypes = [_vp, _d, _d, _d, _i]
        L.sdrpp_cuda_vfo_destroy.argtypes = [_vp, _i]
        L.sdrpp_cuda_vfo_set_offset.argtypes = [_vp, _i, _d]
        L.sdrpp_cuda_vfo_set_bandwidth.argtypes = [_vp, _i, _d]
        L.sdrpp_cuda_vfo_set_out_samplerate.argtypes = [_vp, _i, _d, _d]
        L.sdrpp_cuda_vfo_reset.argtypes = [_vp, _i]
        L.sdrpp_cuda_vfo_info.argtypes = [_vp, _i, _vp]
        L.sdrpp_cuda_frontend_submit.argtypes = [_vp, _i, _vp, _i]
        L.sdrpp_cuda_frontend_submit_device.argtypes = [_vp, _i, _vp, _i]
        L.sdrpp_cuda_frontend_wait.argtypes = [_vp]
        L.sdrpp_cuda_frontend_wait_input.argtypes = [_vp]
        L.sdrpp_cuda_frontend_pending.argtypes = [_vp]
        L.sdrpp_cuda_frontend_drain.argtypes = [_vp]
        L.sdrpp_cuda_frontend_join_streams.argtypes = [_vp]
        L.sdrpp_cuda_comm_unique_id.argtypes = [_vp]
        L.sdrpp_cuda_comm_create.restype = _vp
        L.sdrpp_cuda_comm_create.argtypes = [_vp, _i, _i, _i]
        L.sdrpp_cuda_comm_destroy.argtypes = [_vp]
        L.sdrpp_cuda_comm_info.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp]
        L.sdrpp_cuda_frontend_set_comm.argtypes = [_vp, _vp, _i]
        L.sdrpp_cuda_frontend_submit_shared.argtypes = [_vp, _i, _i]
        L.sdrpp_cuda_frontend_set_fft_display.argtypes = [_vp, _i, C.c_float, _i, C.c_float]
        L.sdrpp_cuda_fft_hold_row.argtypes = [_vp, C.POINTER(_vp)]
        L.sdrpp_cuda_vfo_set_signal_info.argtypes = [_vp, _i, _i]
        L.sdrpp_cuda_frontend_set_snr_smoothing.argtypes = [_vp, _i, C.c_float]
        L.sdrpp_cuda_vfo_signal_info.argtypes = [_vp, _i, _vp, _vp, _vp, _i]
        L.sdrpp_cuda_signal_info.argtypes = [_i, _vp, _i, _vp, _vp, _d, _vp, _vp]
        L.sdrpp_cuda_vfo_output.argtypes = [_vp, _i, C.POINTER(_vp), C.POINTER(_vp)]
        L.sdrpp_cuda_fft_rows.argtypes = [_vp, C.POINTER(_vp)]
        L.sdrpp_cuda_vfo_set_post.argtypes = [_vp, _i, C.POINTER(PostCfg)]
        L.sdrpp_cuda_vfo_set_if_chain.argtypes = [_vp, _i, C.POINTER(IfCfg)]
        L.sdrpp_cuda_vfo_squelch_state.argtypes = [_vp, _i, _vp, _vp]
        L.sdrpp_cuda_vfo_audio.argtypes = [_vp, _i, C.POINTER(_vp)]
        L.sdrpp_cuda_vfo_audio_stereo.argtypes = [_vp, _i, C.POINTER(_vp), C.POINTER(_vp)]
        L.sdrpp_cuda_vfo_rds.argtypes = [_vp, _i, C.POINTER(_vp)]
        L.sdrpp_cuda_frontend_read_iq.argtypes = [_vp, _vp, _i]
        L.sdrpp_cuda_frontend_launches.restype = C.c_longlong
        L.sdrpp_cuda_frontend_launches.argtypes = [_vp]
        L.sdrpp_cuda_frontend_stage1_tensor_launches.restype = C.c_longlong
        L.sdrpp_cuda_frontend_stage1_tensor_launches.argtypes = [_vp]
        L.sdrpp_cuda_frontend_set_stage1_mode.argtypes = [_vp, _i]
        L.sdrpp_cuda_frontend_set_graphs.argtypes = [_vp, _i]
        L.sdrpp_cuda_frontend_graph_stats.argtypes = [_vp, _vp]
        L.sdrpp_cuda_frontend_stream.restype = _vp
        L.sdrpp_cuda_frontend_stream.argtypes = [_vp]
        L.sdrpp_cuda_frontend_kernel_ms.restype = C.c_float
        L.sdrpp_cuda_frontend_kernel_ms.argtypes = [_vp, _i]
        _lib = L
    return _lib


def last_error():
    return lib().sdrpp_cuda_last_error().decode()


def _check(rc, what):
    if rc < 0:
        raise SdrppCudaError(f"{what} failed ({rc}): {last_error()}")
    return rc


def _ptr(a):
    return a.ctypes.data_as(_vp)


def device_count():
    return lib().sdrpp_cuda_device_count()


def init(device=0):
    _check(lib().sdrpp_cuda_init(device), "sdrpp_cuda_init")


# ---- design maths (no GPU) ---------------------------------------------------------------------
def design_window(wtype, size, centered=True):
    buf = np.zeros(size + 2, dtype=np.float32)
    _check(lib().sdrpp_cuda_design_window(wtype, _ptr(buf), size, int(centered)), "design_window")
    return buf[:size].copy()


def design_lowpass(cutoff, trans, sr):
    n = _check(lib().sdrpp_cuda_design_lowpass(cutoff, trans, sr, None, 0), "design_lowpass")
    out = np.zeros(max(n, 1), dtype=np.float32)
    lib().sdrpp_cuda_design_lowpass(cutoff, trans, sr, _ptr(out), n)
    return out[:n]


def design_bandpass_complex(band_start, band_stop, trans, sr, odd=False):
    n = _check(lib().sdrpp_cuda_design_bandpass_complex(band_start, band_stop, trans, sr, int(odd), None, 0), "design_bandpass_complex")
    out = np.zeros(2 * max(n, 1), dtype=np.float32)
    lib().sdrpp_cuda_design_bandpass_complex(band_start, band_stop, trans, sr, int(odd), _ptr(out), n)
    return out[:2 * n].view(np.complex64)


def design_resampler(in_sr, out_sr):
    info = (_i * 6)()
    _check(lib().sdrpp_cuda_design_resampler(in_sr, out_sr, info, None, 0), "design_resampler")
    taps = np.zeros(max(info[4], 1), dtype=np.float32)
    if info[4]:
        lib().sdrpp_cuda_design_resampler(in_sr, out_sr, info, _ptr(taps), info[4])
    return dict(mode=info[0], predec=info[1], interp=info[2], decim=info[3], ntaps=info[4], tpp=info[5]), taps[:info[4]]


def design_decim_plan(ratio):
    dec, cnt, tp = (_i * 4)(), (_i * 4)(), (C.POINTER(C.c_float) * 4)()
    n = lib().sdrpp_cuda_design_decim_plan(ratio, dec, cnt, tp)
    return [(dec[i], np.ctypeslib.as_array(tp[i], shape=(cnt[i],)).copy()) for i in range(n)]


def design_reshape(sr, size, rate):
    skip, nz = _i(), _i()
    lib().sdrpp_cuda_design_reshape(sr, size, rate, C.byref(skip), C.byref(nz))
    return skip.value, nz.value


# ---- one-shot block operations -----------------------------------------------------------------
def _raw(fmt, a):
    a = np.ascontiguousarray(a, dtype=FMT_DTYPE[fmt])
    n = a.size if fmt == FMT_CF32 else (a.size // 6 if fmt == FMT_I24_FILE else a.size // 2)
    return a, n


def convert(fmt, raw):
    raw, n = _raw(fmt, raw)
    out = np.zeros(n, dtype=np.complex64)
    _check(lib().sdrpp_cuda_convert(fmt, _ptr(raw), n, _ptr(out)), "sdrpp_cuda_convert")
    return out


PCM_I8, PCM_I16, PCM_F32 = 0, 1, 2  # dsp/compression/pcm_type.h:4-8


def pcm_decompress(packet):
    """SampleStreamDecompressor::process (sample_stream_decompressor.h:13-36) on one wire packet."""
    packet = np.ascontiguousarray(packet, dtype=np.uint8)
    out = np.zeros(max(1, (len(packet) - 8) // 2), dtype=np.complex64)
    n = _check(lib().sdrpp_cuda_pcm_decompress(_ptr(packet), len(packet), _ptr(out)), "sdrpp_cuda_pcm_decompress")
    return out[:n].copy()


def pcm_compress(pcm_type, x):
    """SampleStreamCompressor::process (sample_stream_compressor.h:26-60): cf32 block -> wire packet."""
    x = np.ascontiguousarray(x, dtype=np.complex64)
    packet = np.zeros(8 + 8 * len(x), dtype=np.uint8)
    n = _check(lib().sdrpp_cuda_pcm_compress(int(pcm_type), _ptr(x), len(x), _ptr(packet)), "sdrpp_cuda_pcm_compress")
    return packet[:n].copy()


def spectrum(N, frame, window, fmt=FMT_CF32, want_X=False):
    frame, n = _raw(fmt, frame)
    window = np.ascontiguousarray(window, dtype=np.float32)
    nz = len(window)
    assert n >= nz
    row = np.zeros(N, dtype=np.float32)
    X = np.zeros(N, dtype=np.complex64) if want_X else None
    _check(lib().sdrpp_cuda_spectrum(N, nz, fmt, _ptr(frame), _ptr(window), _ptr(row), _ptr(X) if want_X else None),
           "sdrpp_cuda_spectrum")
    return (row, X) if want_X else row


def spectrum_device(N, nz, frames, frame_stride, dev_in, window, dev_rows, stream=None):
    """Spectrum rows for frames already in device memory (raw device pointers as ints)."""
    if window is not None:
        window = np.ascontiguousarray(window, dtype=np.float32)
    _check(lib().sdrpp_cuda_spectrum_device(N, nz, frames, frame_stride, C.c_void_p(dev_in), _ptr(window) if window is not None else None, C.c_void_p(dev_rows),
                                            C.c_void_p(stream) if stream else None), "sdrpp_cuda_spectrum_device")


def fft_zoom(row, view_offset, view_bw, whole_bw, out_size):
    """fft_scaler::doZoom on the GPU; returns (pixels, bin boundaries)."""
    row = np.ascontiguousarray(row, dtype=np.float32)
    out = np.zeros(out_size, dtype=np.float32)
    idx = np.zeros(out_size + 1, dtype=np.int32)
    _check(lib().sdrpp_cuda_fft_zoom(len(row), _ptr(row), view_offset, view_bw, whole_bw, out_size, _ptr(out), _ptr(idx)), "sdrpp_cuda_fft_zoom")
    return out, idx


def signal_info(row, center_offsets, bandwidths, whole_bw):
    """WaterFall::calculateVFOSignalInfo on the GPU for several (centerOffset, bandwidth) pairs; returns (strength, snr)."""
    row = np.ascontiguousarray(row, dtype=np.float32)
    co = np.ascontiguousarray(center_offsets, dtype=np.float64)
    bw = np.ascontiguousarray(bandwidths, dtype=np.float64)
    st, sn = np.zeros(len(co), np.float32), np.zeros(len(co), np.float32)
    _check(lib().sdrpp_cuda_signal_info(len(row), _ptr(row), len(co), _ptr(co), _ptr(bw), whole_bw, _ptr(st), _ptr(sn)), "sdrpp_cuda_signal_info")
    return st, sn


# ---- pinned host buffers -----------------------------------------------------------------------
class PinnedArray:
    """A numpy view over pinned host memory (sdrpp_cuda_host_alloc)."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = lib().sdrpp_cuda_host_alloc(max(self.nbytes, 1))
        if not self.ptr:
            raise SdrppCudaError("sdrpp_cuda_host_alloc failed: " + last_error())
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().sdrpp_cuda_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- multi-GPU communicator (one process per GPU; the library owns the NCCL communicator) -----------------------
def comm_unique_id():
    """128 bytes created on one rank (normally 0) and handed to every other rank out of band."""
    buf = (C.c_char * 128)()
    _check(lib().sdrpp_cuda_comm_unique_id(buf), "comm_unique_id")
    return bytes(buf)


class Comm:
    def __init__(self, unique_id, rank, nranks, device):
        assert len(unique_id) == 128
        self.rank, self.nranks = rank, nranks
        self.h = lib().sdrpp_cuda_comm_create(C.c_char_p(unique_id), rank, nranks, device)
        if not self.h:
            raise SdrppCudaError("sdrpp_cuda_comm_create failed: " + last_error())

    def info(self):
        r, n, v = _i(), _i(), _i()
        b, by = C.c_longlong(), C.c_longlong()
        _check(lib().sdrpp_cuda_comm_info(self.h, C.byref(r), C.byref(n), C.byref(v), C.byref(b), C.byref(by)), "comm_info")
        return dict(rank=r.value, nranks=n.value, nccl_version=v.value, broadcasts=b.value, bytes=by.value)

    def close(self):
        if getattr(self, "h", None):
            lib().sdrpp_cuda_comm_destroy(self.h)
            self.h = None


# ---- front end ---------------------------------------------------------------------------------
class Frontend:
    """sigpath::iqFrontEnd + the VFO set of sigpath::vfoManager on one GPU."""

    def __init__(self, sample_rate, decim_ratio=1, dc_blocking=False, invert_iq=False, fft_size=0, fft_rate=20.0,
                 fft_window=WIN_NUTTALL, max_block=1000000, ring_log2=0, max_fft_rows=0):
        cfg = FrontendCfg(sample_rate, decim_ratio, int(dc_blocking), int(invert_iq), fft_size, fft_rate, fft_window,
                          max_block, ring_log2, max_fft_rows)
        self.fft_size = fft_size
        self.h = lib().sdrpp_cuda_frontend_create(C.byref(cfg))
        if not self.h:
            raise SdrppCudaError("sdrpp_cuda_frontend_create failed: " + last_error())
        self.demod = {}

    def close(self):
        if getattr(self, "h", None):
            lib().sdrpp_cuda_frontend_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # setters
    def set_sample_rate(self, sr): _check(lib().sdrpp_cuda_frontend_set_sample_rate(self.h, sr), "set_sample_rate")
    def set_decimation(self, r): _check(lib().sdrpp_cuda_frontend_set_decimation(self.h, r), "set_decimation")
    def set_dc_blocking(self, e): _check(lib().sdrpp_cuda_frontend_set_dc_blocking(self.h, int(e)), "set_dc_blocking")
    def set_invert_iq(self, e): _check(lib().sdrpp_cuda_frontend_set_invert_iq(self.h, int(e)), "set_invert_iq")
    def set_fft_window(self, w): _check(lib().sdrpp_cuda_frontend_set_fft_window(self.h, w), "set_fft_window")
    def set_fft_rate(self, r): _check(lib().sdrpp_cuda_frontend_set_fft_rate(self.h, r), "set_fft_rate")
    def set_readback(self, e): _check(lib().sdrpp_cuda_frontend_set_readback(self.h, int(e)), "set_readback")
    def set_profiling(self, e): _check(lib().sdrpp_cuda_frontend_set_profiling(self.h, int(e)), "set_profiling")

    def set_fft_size(self, n):
        _check(lib().sdrpp_cuda_frontend_set_fft_size(self.h, n), "set_fft_size")
        self.fft_size = n

    @property
    def effective_samplerate(self):
        return lib().sdrpp_cuda_frontend_effective_samplerate(self.h)

    # VFOs
    def add_vfo(self, out_sr, bw, offset, demod=DEMOD_NONE):
        vid = _check(lib().sdrpp_cuda_vfo_create(self.h, out_sr, bw, offset, demod), "vfo_create")
        self.demod[vid] = demod
        return vid

    def remove_vfo(self, vid):
        _check(lib().sdrpp_cuda_vfo_destroy(self.h, vid), "vfo_destroy")
        self.demod.pop(vid, None)

    def vfo_set_offset(self, vid, off): _check(lib().sdrpp_cuda_vfo_set_offset(self.h, vid, off), "vfo_set_offset")
    def vfo_set_bandwidth(self, vid, bw): _check(lib().sdrpp_cuda_vfo_set_bandwidth(self.h, vid, bw), "vfo_set_bandwidth")
    def vfo_set_out_samplerate(self, vid, sr, bw): _check(lib().sdrpp_cuda_vfo_set_out_samplerate(self.h, vid, sr, bw), "vfo_set_out_samplerate")
    def vfo_reset(self, vid): _check(lib().sdrpp_cuda_vfo_reset(self.h, vid), "vfo_reset")

    def vfo_info(self, vid):
        a = (_i * 9)()
        _check(lib().sdrpp_cuda_vfo_info(self.h, vid, a), "vfo_info")
        return dict(mode=a[0], predec=a[1], interp=a[2], decim=a[3], rtaps=a[4], tpp=a[5], ftaps=a[6], s1_decim=a[7], s1_taps=a[8])

    # blocks
    def submit(self, fmt, raw, count=None):
        """raw: numpy array (any host memory) or PinnedArray."""
        if isinstance(raw, PinnedArray):
            n = count if count is not None else raw.nbytes // FMT_BYTES[fmt]
            _check(lib().sdrpp_cuda_frontend_submit(self.h, fmt, raw.ptr, n), "submit")
            return n
        raw, n = _raw(fmt, raw)
        if count is not None:
            n = count
        self._keep = raw
        _check(lib().sdrpp_cuda_frontend_submit(self.h, fmt, _ptr(raw), n), "submit")
        return n

    def submit_pcm(self, packet):
        """One SDR++ server wire packet (host memory) through the whole path; returns the sample count."""
        packet = np.ascontiguousarray(packet, dtype=np.uint8)
        self._keep = packet
        return _check(lib().sdrpp_cuda_frontend_submit_pcm(self.h, _ptr(packet), len(packet)), "submit_pcm")

    def submit_device(self, fmt, dev_ptr, count):
        _check(lib().sdrpp_cuda_frontend_submit_device(self.h, fmt, C.c_void_p(dev_ptr), count), "submit_device")

    def wait(self):
        _check(lib().sdrpp_cuda_frontend_wait(self.h), "wait")

    def set_comm(self, comm, root=0):
        """Attach to a communicator: submits on `root` broadcast the raw block, the other ranks call submit_shared."""
        _check(lib().sdrpp_cuda_frontend_set_comm(self.h, comm.h if comm is not None else None, root), "set_comm")

    def submit_shared(self, fmt, count):
        _check(lib().sdrpp_cuda_frontend_submit_shared(self.h, fmt, count), "submit_shared")

    def join_streams(self):
        _check(lib().sdrpp_cuda_frontend_join_streams(self.h), "join_streams")

    def wait_input(self):
        _check(lib().sdrpp_cuda_frontend_wait_input(self.h), "wait_input")

    def drain(self):
        """Wait for everything in flight and discard results not yet waited for."""
        _check(lib().sdrpp_cuda_frontend_drain(self.h), "drain")

    @property
    def pending(self):
        return lib().sdrpp_cuda_frontend_pending(self.h)

    def process(self, fmt, raw):
        self.submit(fmt, raw)
        self.wait()

    def vfo_output(self, vid, copy=True):
        iq, dm = _vp(), _vp()
        n = _check(lib().sdrpp_cuda_vfo_output(self.h, vid, C.byref(iq), C.byref(dm)), "vfo_output")
        if n == 0:
            return np.zeros(0, np.complex64), (np.zeros(0, np.float32) if self.demod.get(vid) else None)
        a = np.ctypeslib.as_array(C.cast(iq, C.POINTER(C.c_float)), shape=(2 * n,)).view(np.complex64)
        d = None
        if dm.value:
            d = np.ctypeslib.as_array(C.cast(dm, C.POINTER(C.c_float)), shape=(n,))
        return (a.copy(), d.copy() if d is not None else None) if copy else (a, d)

    def set_post(self, vid, enabled=True, fm_lowpass=True, am_agc_mode=0, ssb_agc=True, agc_attack=0.0, agc_decay=0.0,
                 dc_block_rate=0.0, agc_gain=0.0, wfm=False, wfm_stereo=True, wfm_rds=False):
        """Post-detector stages of the demodulator behind the VFO (dsp::demod::FM / AM / SSB; wfm: dsp::demod::BroadcastFM)."""
        cfg = PostCfg(int(enabled), int(fm_lowpass), int(am_agc_mode), int(ssb_agc), agc_attack, agc_decay, dc_block_rate, agc_gain, int(wfm), int(wfm_stereo), int(wfm_rds))
        _check(lib().sdrpp_cuda_vfo_set_post(self.h, vid, C.byref(cfg)), "vfo_set_post")

    def set_if_chain(self, vid, nb=False, nb_rate=500.0 / 24000.0, nb_level=10.0, squelch=False, squelch_level=-100.0, fmif_bins=0):
        """Radio IF chain in front of the demodulator: NoiseBlanker -> Squelch -> FMIF (radio_module.h:73-78)."""
        cfg = IfCfg(int(nb), nb_rate, nb_level, int(squelch), squelch_level, int(fmif_bins))
        _check(lib().sdrpp_cuda_vfo_set_if_chain(self.h, vid, C.byref(cfg)), "vfo_set_if_chain")

    def squelch_state(self, vid):
        m, l = C.c_int(0), C.c_float(0)
        _check(lib().sdrpp_cuda_vfo_squelch_state(self.h, vid, C.byref(m), C.byref(l)), "vfo_squelch_state")
        return bool(m.value), float(l.value)

    def vfo_audio(self, vid):
        p = _vp()
        n = _check(lib().sdrpp_cuda_vfo_audio(self.h, vid, C.byref(p)), "vfo_audio")
        if n == 0:
            return np.zeros(0, np.float32)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,)).copy()

    def vfo_audio_stereo(self, vid):
        l, r = _vp(), _vp()
        n = _check(lib().sdrpp_cuda_vfo_audio_stereo(self.h, vid, C.byref(l), C.byref(r)), "vfo_audio_stereo")
        if n == 0:
            return np.zeros(0, np.float32), np.zeros(0, np.float32)
        return (np.ctypeslib.as_array(C.cast(l, C.POINTER(C.c_float)), shape=(n,)).copy(),
                np.ctypeslib.as_array(C.cast(r, C.POINTER(C.c_float)), shape=(n,)).copy())

    def vfo_rds(self, vid):
        """BroadcastFM::rdsOut of the last waited block: complex samples at 5 kS/s (needs set_post(wfm=True, wfm_rds=True))."""
        p = _vp()
        n = _check(lib().sdrpp_cuda_vfo_rds(self.h, vid, C.byref(p)), "vfo_rds")
        if n == 0:
            return np.zeros(0, np.complex64)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(2 * n,)).copy().view(np.complex64)

    def fft_rows(self, copy=True):
        p = _vp()
        n = _check(lib().sdrpp_cuda_fft_rows(self.h, C.byref(p)), "fft_rows")
        if n == 0:
            return np.zeros((0, self.fft_size), np.float32)
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n, self.fft_size))
        return a.copy() if copy else a

    def set_fft_zoom(self, view_offset, view_bw, whole_bw, out_size, keep_raw=True):
        _check(lib().sdrpp_cuda_frontend_set_fft_zoom(self.h, view_offset, view_bw, whole_bw, out_size, int(keep_raw)), "set_fft_zoom")
        self.zoom_out = out_size

    def fft_zoomed_rows(self, copy=True):
        p = _vp()
        n = _check(lib().sdrpp_cuda_fft_zoomed_rows(self.h, C.byref(p)), "fft_zoomed_rows")
        if n == 0:
            return np.zeros((0, getattr(self, "zoom_out", 0)), np.float32)
        a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n, self.zoom_out))
        return a.copy() if copy else a

    def set_fft_display(self, smoothing=False, smoothing_speed=0.5, hold=False, hold_speed=0.3):
        """FFT smoothing and peak hold on the zoomed row (WaterFall::pushFFT)."""
        _check(lib().sdrpp_cuda_frontend_set_fft_display(self.h, int(smoothing), smoothing_speed, int(hold), hold_speed), "set_fft_display")

    def fft_hold_row(self):
        p = _vp()
        n = _check(lib().sdrpp_cuda_fft_hold_row(self.h, C.byref(p)), "fft_hold_row")
        if n == 0:
            return np.zeros(0, np.float32)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_float)), shape=(n,)).copy()

    def vfo_set_signal_info(self, vid, enabled=True):
        _check(lib().sdrpp_cuda_vfo_set_signal_info(self.h, vid, int(enabled)), "vfo_set_signal_info")

    def set_snr_smoothing(self, enabled, speed=0.5):
        _check(lib().sdrpp_cuda_frontend_set_snr_smoothing(self.h, int(enabled), speed), "set_snr_smoothing")

    def vfo_signal_info(self, vid, cap=64):
        st, sn = np.zeros(cap, np.float32), np.zeros(cap, np.float32)
        mx = C.c_float(0)
        n = _check(lib().sdrpp_cuda_vfo_signal_info(self.h, vid, _ptr(st), _ptr(sn), C.byref(mx), cap), "vfo_signal_info")
        n = min(n, cap)
        return st[:n], sn[:n], float(mx.value)

    def read_iq(self, cap):
        out = np.zeros(cap, dtype=np.complex64)
        n = _check(lib().sdrpp_cuda_frontend_read_iq(self.h, _ptr(out), cap), "read_iq")
        return out[:n]

    @property
    def launches(self):
        return lib().sdrpp_cuda_frontend_launches(self.h)

    @property
    def stage1_tensor_launches(self):
        return lib().sdrpp_cuda_frontend_stage1_tensor_launches(self.h)

    def set_graphs(self, enabled):
        """CUDA-graph replay of repeated blocks (default on); off: every block command by command. Same results."""
        _check(lib().sdrpp_cuda_frontend_set_graphs(self.h, int(enabled)), "set_graphs")

    def graph_stats(self):
        out = (C.c_longlong * 4)()
        _check(lib().sdrpp_cuda_frontend_graph_stats(self.h, out), "graph_stats")
        return {"replayed_runs": int(out[0]), "graphs_instantiated": int(out[1]), "direct_runs": int(out[2]), "instantiate_us": int(out[3])}

    def set_stage1_mode(self, mode):
        """0: tensor cores where the plan allows (default), 1: FP32 FMA kernel only."""
        _check(lib().sdrpp_cuda_frontend_set_stage1_mode(self.h, int(mode)), "set_stage1_mode")

    @property
    def stream(self):
        return lib().sdrpp_cuda_frontend_stream(self.h)

    def kernel_ms(self):
        return [lib().sdrpp_cuda_frontend_kernel_ms(self.h, i) for i in range(4)]
