// TEST INFRASTRUCTURE ONLY (oracle). Never linked into or called by the product path.
//
// oracle/_ref/libsdrpp_ref*.so: the reference's OWN dsp/ headers, included read-only from
// /root/reference/core/src (never copied into this repo), compiled against the generic-VOLK shim
// in oracle/shim/, and exposed through a flat C API so Python tests can drive them with ctypes.
// Every entry point below calls the reference class named in its comment; nothing here restates
// DSP arithmetic except where the reference code cannot be compiled in this image:
//   * IQFrontEnd::handler (signal_path/iq_frontend.cpp:230-249) drags in the GUI and FFTW, so the
//     three-line window*x -> DFT -> power-spectrum body is restated in ref_spectrum() around an
//     in-file FFT (FFTW3f is not installed; a forward unnormalised DFT is what it computes).
// Build: see oracle/Makefile (g++ -std=c++17 -Ioracle/shim -I/root/reference/core/src).
#include <dsp/channel/rx_vfo.h>
#include <dsp/channel/frequency_xlator.h>
#include <dsp/multirate/power_decimator.h>
#include <dsp/multirate/polyphase_resampler.h>
#include <dsp/multirate/rational_resampler.h>
#include <dsp/filter/fir.h>
#include <dsp/filter/decimating_fir.h>
#include <dsp/taps/low_pass.h>
#include <dsp/taps/from_array.h>
#include <dsp/window/window.h>
#include <dsp/correction/dc_blocker.h>
#include <dsp/demod/quadrature.h>
#include <dsp/math/conjugate.h>
#include <dsp/convert/complex_to_real.h>
#include <gui/widgets/fft_scaler.h>
#include <dsp/demod/fm.h>
#include <dsp/demod/am.h>
#include <dsp/demod/ssb.h>
#include <dsp/demod/broadcast_fm.h>
#include <dsp/compression/sample_stream_compressor.h>
#include <dsp/noise_reduction/noise_blanker.h>
#include <dsp/noise_reduction/squelch.h>
#include <dsp/noise_reduction/fm_if.h>
#include <dsp/compression/sample_stream_decompressor.h>

#include <vector>
#include <thread>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>

using dsp::complex_t;

#define API extern "C" __attribute__((visibility("default")))

// ---------------------------------------------------------------------------------------------
// Accessors for protected state (no edits to the reference: derive + re-export)
// ---------------------------------------------------------------------------------------------
struct ResampPeek : dsp::multirate::RationalResampler<complex_t> {
    using Base = dsp::multirate::RationalResampler<complex_t>;
    using Base::mode; using Base::decim; using Base::resamp; using Base::rtaps;
};
struct PolyPeek : dsp::multirate::PolyphaseResampler<complex_t> {
    using Base = dsp::multirate::PolyphaseResampler<complex_t>;
    using Base::_interp; using Base::_decim; using Base::phase; using Base::offset; using Base::phases;
};
struct PDecPeek : dsp::multirate::PowerDecimator<complex_t> {
    using Base = dsp::multirate::PowerDecimator<complex_t>;
    using Base::decimFirs; using Base::_ratio; using Base::stageCount;
};
struct DFirPeek : dsp::filter::DecimatingFIR<complex_t, float> {
    using Base = dsp::filter::DecimatingFIR<complex_t, float>;
    using Base::offset; using Base::_decimation; using Base::_taps;
};
struct VfoPeek : dsp::channel::RxVFO {
    using Base = dsp::channel::RxVFO;
    using Base::xlator; using Base::resamp; using Base::filter; using Base::ftaps; using Base::filterNeeded;
};
struct XlatPeek : dsp::channel::FrequencyXlator {
    using Base = dsp::channel::FrequencyXlator;
    using Base::phase; using Base::phaseDelta;
};

// ---------------------------------------------------------------------------------------------
// Window / taps / plans (host-side design maths, all double -> float in the reference)
// ---------------------------------------------------------------------------------------------
// dsp::window::createWindow (dsp/window/window.h:38-64). buf must hold size+1 floats (the centred
// branch writes buffer[i+1] for odd sizes).
API void ref_create_window(int type, float* buf, int size, int centered) {
    dsp::window::createWindow((dsp::window::windowType)type, buf, size, centered != 0);
}

// dsp::taps::lowPass (dsp/taps/low_pass.h:7-11). Returns the tap count; copies min(count,cap).
API int ref_lowpass_taps(double cutoff, double transWidth, double sampleRate, float* out, int cap) {
    dsp::tap<float> t = dsp::taps::lowPass(cutoff, transWidth, sampleRate);
    int n = (int)t.size;
    if (out) { memcpy(out, t.taps, sizeof(float) * std::min(n, cap)); }
    dsp::taps::free(t);
    return n;
}

// dsp::multirate::decim::plans (dsp/multirate/decim/plans.h:126-140). ratio = 2^k.
// Writes up to 4 stages; returns the stage count (0 if the ratio is invalid).
API int ref_decim_plan(int ratio, int* decimation, int* tapcount, const float** taps) {
    if (ratio < 2 || (ratio & (ratio - 1))) return 0;
    int id = (int)log2((double)ratio) - 1;
    if (id < 0 || id >= (int)dsp::multirate::decim::plans_len) return 0;
    const dsp::multirate::decim::plan& p = dsp::multirate::decim::plans[id];
    for (unsigned i = 0; i < p.stageCount; i++) {
        decimation[i] = (int)p.stages[i].decimation;
        tapcount[i] = (int)p.stages[i].tapcount;
        taps[i] = p.stages[i].taps;
    }
    return (int)p.stageCount;
}

// ---------------------------------------------------------------------------------------------
// Block objects: create / process / reset / destroy
// ---------------------------------------------------------------------------------------------
struct FirObj { dsp::tap<float> taps; dsp::filter::FIR<complex_t, float> fir; };
// dsp::filter::FIR<complex_t,float> (dsp/filter/fir.h)
API void* ref_fir_create(const float* taps, int n) {
    FirObj* o = new FirObj;
    o->taps = dsp::taps::fromArray<float>(n, taps);
    o->fir.init(NULL, o->taps);
    return o;
}
API int ref_fir_process(void* h, int count, const complex_t* in, complex_t* out) { return ((FirObj*)h)->fir.process(count, in, out); }
API void ref_fir_reset(void* h) { ((FirObj*)h)->fir.reset(); }
API void ref_fir_destroy(void* h) { FirObj* o = (FirObj*)h; delete o; }

struct DFirObj { dsp::tap<float> taps; dsp::filter::DecimatingFIR<complex_t, float> fir; };
// dsp::filter::DecimatingFIR<complex_t,float> (dsp/filter/decimating_fir.h)
API void* ref_decfir_create(const float* taps, int n, int decim) {
    DFirObj* o = new DFirObj;
    o->taps = dsp::taps::fromArray<float>(n, taps);
    o->fir.init(NULL, o->taps, decim);
    return o;
}
API int ref_decfir_process(void* h, int count, const complex_t* in, complex_t* out) { return ((DFirObj*)h)->fir.process(count, in, out); }
API int ref_decfir_offset(void* h) { return static_cast<DFirPeek*>(&((DFirObj*)h)->fir)->offset; }
API void ref_decfir_reset(void* h) { ((DFirObj*)h)->fir.reset(); }
API void ref_decfir_destroy(void* h) { delete (DFirObj*)h; }

// dsp::multirate::PowerDecimator<complex_t> (dsp/multirate/power_decimator.h)
API void* ref_powerdecim_create(int ratio) { return new dsp::multirate::PowerDecimator<complex_t>(NULL, (unsigned)ratio); }
API int ref_powerdecim_process(void* h, int count, const complex_t* in, complex_t* out) {
    return ((dsp::multirate::PowerDecimator<complex_t>*)h)->process(count, in, out);
}
API int ref_powerdecim_offsets(void* h, int* offsets) {
    PDecPeek* p = static_cast<PDecPeek*>((dsp::multirate::PowerDecimator<complex_t>*)h);
    if (p->_ratio == 1) return 0;
    for (int i = 0; i < p->stageCount; i++) offsets[i] = static_cast<DFirPeek*>(p->decimFirs[i])->offset;
    return p->stageCount;
}
API void ref_powerdecim_reset(void* h) { ((dsp::multirate::PowerDecimator<complex_t>*)h)->reset(); }
API void ref_powerdecim_destroy(void* h) { delete (dsp::multirate::PowerDecimator<complex_t>*)h; }

struct PolyObj { dsp::tap<float> taps; dsp::multirate::PolyphaseResampler<complex_t> r; };
// dsp::multirate::PolyphaseResampler<complex_t> (dsp/multirate/polyphase_resampler.h)
API void* ref_polyphase_create(int interp, int decim, const float* taps, int n) {
    PolyObj* o = new PolyObj;
    o->taps = dsp::taps::fromArray<float>(n, taps);
    o->r.init(NULL, interp, decim, o->taps);
    return o;
}
API int ref_polyphase_process(void* h, int count, const complex_t* in, complex_t* out) { return ((PolyObj*)h)->r.process(count, in, out); }
API void ref_polyphase_state(void* h, int* phase, int* offset) {
    PolyPeek* p = static_cast<PolyPeek*>(&((PolyObj*)h)->r);
    *phase = p->phase; *offset = p->offset;
}
API void ref_polyphase_destroy(void* h) { delete (PolyObj*)h; }

// dsp::multirate::RationalResampler<complex_t> (dsp/multirate/rational_resampler.h)
API void* ref_resampler_create(double inSR, double outSR) { return new dsp::multirate::RationalResampler<complex_t>(NULL, inSR, outSR); }
API int ref_resampler_process(void* h, int count, const complex_t* in, complex_t* out) {
    return ((dsp::multirate::RationalResampler<complex_t>*)h)->process(count, in, out);
}
// info[0]=mode (0 BOTH,1 DECIM_ONLY,2 RESAMP_ONLY,3 NONE) [1]=predec ratio [2]=interp [3]=decim
// [4]=resampler tap count [5]=taps per phase. Copies the (already interp-scaled) taps if asked.
static void resampler_info(ResampPeek* r, int* info, float* taps, int cap) {
    const int mode = (int)r->mode;   // enum order: BOTH, DECIM_ONLY, RESAMP_ONLY, NONE (rational_resampler.h:114-119)
    info[0] = mode;
    PDecPeek* pd = static_cast<PDecPeek*>(&r->decim);
    PolyPeek* pp = static_cast<PolyPeek*>(&r->resamp);
    bool useDecim = (mode == 0 || mode == 1);
    bool usePoly = (mode == 0 || mode == 2);
    info[1] = useDecim ? (int)pd->_ratio : 1;
    info[2] = usePoly ? pp->_interp : 1;
    info[3] = usePoly ? pp->_decim : 1;
    info[4] = usePoly ? (int)r->rtaps.size : 0;
    info[5] = usePoly ? pp->phases.tapsPerPhase : 0;
    if (taps && usePoly) memcpy(taps, r->rtaps.taps, sizeof(float) * std::min<int>(r->rtaps.size, cap));
}
API void ref_resampler_info(void* h, int* info, float* taps, int cap) {
    resampler_info(static_cast<ResampPeek*>((dsp::multirate::RationalResampler<complex_t>*)h), info, taps, cap);
}
API void ref_resampler_destroy(void* h) { delete (dsp::multirate::RationalResampler<complex_t>*)h; }

// dsp::channel::FrequencyXlator (dsp/channel/frequency_xlator.h)
API void* ref_xlator_create(double offsetHz, double sampleRate) { return new dsp::channel::FrequencyXlator(NULL, offsetHz, sampleRate); }
API int ref_xlator_process(void* h, int count, const complex_t* in, complex_t* out) { return ((dsp::channel::FrequencyXlator*)h)->process(count, in, out); }
API void ref_xlator_set_offset(void* h, double offsetHz, double sampleRate) { ((dsp::channel::FrequencyXlator*)h)->setOffset(offsetHz, sampleRate); }
API void ref_xlator_state(void* h, float* phase, float* delta) {
    XlatPeek* x = static_cast<XlatPeek*>((dsp::channel::FrequencyXlator*)h);
    phase[0] = x->phase.real(); phase[1] = x->phase.imag();
    delta[0] = x->phaseDelta.real(); delta[1] = x->phaseDelta.imag();
}
API void ref_xlator_destroy(void* h) { delete (dsp::channel::FrequencyXlator*)h; }

// "Ideal NCO" flavour: the reference's FrequencyXlator with its fp32 phase recurrence replaced by the closed
// form n * arg(phaseDelta) in extended precision -- the SAME fp32-quantised increment the reference computes in
// FrequencyXlator::init (frequency_xlator.h:15-23), the same fp32 complex multiply per sample, everything downstream the
// reference's own blocks. |ref_f32 - ideal| is the reference rotator's own random walk, |gpu - ideal| the GPU's error.
struct IdealNco {
    long double turns = 0, phi = 0;
    void setOffset(double offsetHz, double sampleRate) {
        dsp::channel::FrequencyXlator x(NULL, offsetHz, sampleRate);        // the reference rounds the increment
        XlatPeek* p = static_cast<XlatPeek*>(&x);
        turns = (long double)atan2((double)p->phaseDelta.imag(), (double)p->phaseDelta.real()) / (2.0L * 3.14159265358979323846264338327950288L);
    }
    void reset() { phi = 0; }
    void process(int count, const complex_t* in, complex_t* out) {
        for (int i = 0; i < count; i++) {
            const double a = 2.0 * M_PI * (double)phi;
            const lv_32fc_t ph((float)cos(a), (float)sin(a));
            const lv_32fc_t y = oracle_cmul(lv_32fc_t(in[i].re, in[i].im), ph); // the rotator's per-sample product (volk shim)
            out[i].re = y.real(); out[i].im = y.imag();
            phi += turns;
            phi -= floorl(phi + 0.5L);
        }
    }
};

// dsp::channel::RxVFO (dsp/channel/rx_vfo.h). ideal: RxVFO at offset 0 (its xlator multiplies by exactly 1) behind an IdealNco.
struct VfoObj {
    dsp::channel::RxVFO* vfo = nullptr;
    bool ideal = false;
    IdealNco nco;
    double inSR = 0;
    std::vector<complex_t> tmp;
    ~VfoObj() { delete vfo; }
};
API void* ref_rxvfo_create(double inSR, double outSR, double bw, double offset) {
    VfoObj* o = new VfoObj;
    o->vfo = new dsp::channel::RxVFO(NULL, inSR, outSR, bw, offset);
    o->inSR = inSR;
    return o;
}
API void* ref_rxvfo_create_ideal(double inSR, double outSR, double bw, double offset) {
    VfoObj* o = new VfoObj;
    o->vfo = new dsp::channel::RxVFO(NULL, inSR, outSR, bw, 0.0);
    o->ideal = true; o->inSR = inSR;
    o->nco.setOffset(-offset, inSR);                                        // xlator.init(NULL, -_offset, _inSamplerate), rx_vfo.h:27
    return o;
}
API int ref_rxvfo_process(void* h, int count, const complex_t* in, complex_t* out) {
    VfoObj* o = (VfoObj*)h;
    if (!o->ideal) return o->vfo->process(count, in, out);
    o->tmp.resize(count + 16);
    o->nco.process(count, in, o->tmp.data());
    return o->vfo->process(count, o->tmp.data(), out);
}
API void ref_rxvfo_set_offset(void* h, double offset) {
    VfoObj* o = (VfoObj*)h;
    if (o->ideal) o->nco.setOffset(-offset, o->inSR); else o->vfo->setOffset(offset);
}
API void ref_rxvfo_set_bandwidth(void* h, double bw) { ((VfoObj*)h)->vfo->setBandwidth(bw); }
API void ref_rxvfo_set_out_samplerate(void* h, double outSR, double bw) { ((VfoObj*)h)->vfo->setOutSamplerate(outSR, bw); }
API void ref_rxvfo_set_in_samplerate(void* h, double inSR) { ((VfoObj*)h)->vfo->setInSamplerate(inSR); ((VfoObj*)h)->inSR = inSR; }
API void ref_rxvfo_reset(void* h) { VfoObj* o = (VfoObj*)h; o->vfo->reset(); o->nco.reset(); }
// info[0..5] as ref_resampler_info, info[6]=channel filter tap count (0 when bypassed).
API void ref_rxvfo_info(void* h, int* info, float* rtaps, int rcap, float* ftaps, int fcap) {
    VfoPeek* v = static_cast<VfoPeek*>(((VfoObj*)h)->vfo);
    resampler_info(static_cast<ResampPeek*>(&v->resamp), info, rtaps, rcap);
    info[6] = v->filterNeeded ? (int)v->ftaps.size : 0;
    if (ftaps && v->filterNeeded) memcpy(ftaps, v->ftaps.taps, sizeof(float) * std::min<int>(v->ftaps.size, fcap));
}
API void ref_rxvfo_destroy(void* h) { delete (VfoObj*)h; }

// FrequencyXlator alone, ideal flavour
API void* ref_xlator_create_ideal(double offsetHz, double sampleRate) { IdealNco* n = new IdealNco; n->setOffset(offsetHz, sampleRate); return n; }
API int ref_xlatorideal_process(void* h, int count, const complex_t* in, complex_t* out) { ((IdealNco*)h)->process(count, in, out); return count; }
API void ref_xlatorideal_destroy(void* h) { delete (IdealNco*)h; }

// dsp::correction::DCBlocker<complex_t> (dsp/correction/dc_blocker.h); rate = 50/effectiveSr in
// IQFrontEnd::genDCBlockRate (signal_path/iq_frontend.h:52-54)
API void* ref_dcblock_create(double rate) { return new dsp::correction::DCBlocker<complex_t>(NULL, rate); }
API int ref_dcblock_process(void* h, int count, complex_t* in, complex_t* out) { return ((dsp::correction::DCBlocker<complex_t>*)h)->process(count, in, out); }
API void ref_dcblock_destroy(void* h) { delete (dsp::correction::DCBlocker<complex_t>*)h; }

// dsp::math::Conjugate::process (dsp/math/conjugate.h:12-15)
API int ref_conjugate(int count, const complex_t* in, complex_t* out) { return dsp::math::Conjugate::process(count, in, out); }

// dsp::demod::Quadrature (dsp/demod/quadrature.h). reset() is called at creation because _din is
// otherwise uninitialised.
API void* ref_quadrature_create(double deviation, double sampleRate) {
    auto* q = new dsp::demod::Quadrature(NULL, deviation, sampleRate);
    q->reset();
    return q;
}
API int ref_quadrature_process(void* h, int count, complex_t* in, float* out) { return ((dsp::demod::Quadrature*)h)->process(count, in, out); }
API void ref_quadrature_destroy(void* h) { delete (dsp::demod::Quadrature*)h; }

// AM front end: volk_32fc_magnitude_32f as called by dsp::demod::AM::process (dsp/demod/am.h:122)
API int ref_am_magnitude(int count, const complex_t* in, float* out) {
    volk_32fc_magnitude_32f(out, (const lv_32fc_t*)in, count);
    return count;
}

// SSB front end: FrequencyXlator at getTranslation() then ComplexToReal, as in
// dsp::demod::SSB::process (dsp/demod/ssb.h:90-101, translation :119-126). mode 0 USB, 1 LSB, 2 DSB.
struct SsbObj { dsp::channel::FrequencyXlator x; std::vector<complex_t> tmp; bool ideal = false; IdealNco nco; };
API void* ref_ssb_create(int mode, double bandwidth, double sampleRate) {
    SsbObj* o = new SsbObj;
    double tr = (mode == 0) ? bandwidth / 2.0 : (mode == 1) ? -bandwidth / 2.0 : 0.0;
    o->x.init(NULL, tr, sampleRate);
    return o;
}
API void* ref_ssb_create_ideal(int mode, double bandwidth, double sampleRate) {
    SsbObj* o = new SsbObj;
    double tr = (mode == 0) ? bandwidth / 2.0 : (mode == 1) ? -bandwidth / 2.0 : 0.0;
    o->x.init(NULL, tr, sampleRate);
    o->ideal = true; o->nco.setOffset(tr, sampleRate);
    return o;
}
API int ref_ssb_process(void* h, int count, const complex_t* in, float* out) {
    SsbObj* o = (SsbObj*)h;
    o->tmp.resize(count);
    if (o->ideal) o->nco.process(count, in, o->tmp.data());
    else o->x.process(count, in, o->tmp.data());
    return dsp::convert::ComplexToReal::process(count, o->tmp.data(), out);
}
API void ref_ssb_destroy(void* h) { delete (SsbObj*)h; }

// ---------------------------------------------------------------------------------------------
// Spectrum line: restatement of IQFrontEnd::handler + updateFFTSize
// (signal_path/iq_frontend.cpp:230-249,272-296). Power-of-two sizes only (the GUI list,
// gui/menus/display.cpp:33-45).
// ---------------------------------------------------------------------------------------------
template <class R>
static void fft_pow2(std::vector<std::complex<R>>& a) {
    const size_t n = a.size();
    // bit reversal
    for (size_t i = 1, j = 0; i < n; i++) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    // twiddles in double regardless of R, rounded once
    std::vector<std::complex<R>> w(n / 2);
    for (size_t k = 0; k < n / 2; k++) {
        double ang = -2.0 * M_PI * (double)k / (double)n;
        w[k] = std::complex<R>((R)cos(ang), (R)sin(ang));
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        const size_t half = len >> 1, step = n / len;
        for (size_t i = 0; i < n; i += len) {
            for (size_t k = 0; k < half; k++) {
                std::complex<R> u = a[i + k];
                std::complex<R> t = a[i + k + half];
                const std::complex<R> tw = w[k * step];
                std::complex<R> v(t.real() * tw.real() - t.imag() * tw.imag(), t.real() * tw.imag() + t.imag() * tw.real());
                a[i + k] = u + v;
                a[i + k + half] = u - v;
            }
        }
    }
}

// frame: nz samples; window: nz floats (from ref_create_window(..., centered=1)).
// X64 (optional): N complex doubles, the fp64 DFT of the fp32 windowed frame ("ref_f64").
// row32: N floats computed the reference way (fp32 FFT + VOLK power spectrum) ("ref_f32").
// row64 (optional): N doubles, 10*log10 |X64|^2.
API int ref_spectrum(int N, int nz, const complex_t* frame, const float* window,
                     float* row32, double* X64, double* row64) {
    if (N <= 0 || (N & (N - 1)) || nz > N) return -1;
    std::vector<lv_32fc_t> fftIn(N, lv_32fc_t(0, 0));
    volk_32fc_32f_multiply_32fc(fftIn.data(), (const lv_32fc_t*)frame, window, nz);
    if (row32) {
        std::vector<std::complex<float>> a(fftIn.begin(), fftIn.end());
        fft_pow2<float>(a);
        volk_32fc_s32f_power_spectrum_32f(row32, a.data(), 1.0f, N);
    }
    if (X64 || row64) {
        std::vector<std::complex<double>> a(N);
        for (int i = 0; i < N; i++) a[i] = std::complex<double>(fftIn[i].real(), fftIn[i].imag());
        fft_pow2<double>(a);
        for (int i = 0; i < N; i++) {
            if (X64) { X64[2 * i] = a[i].real(); X64[2 * i + 1] = a[i].imag(); }
            if (row64) { row64[i] = 10.0 * log10(a[i].real() * a[i].real() + a[i].imag() * a[i].imag()); }
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// CPU baseline harness, shaped like dsp::bench::SpeedTester (dsp/bench/speed_tester.h:31-56) and
// the reference's thread-per-block model (dsp/block.h:70-76): one worker per VFO calling
// RxVFO::process() + the demod front end on the common block; workers are multiplexed onto
// `nthreads` OS threads. An optional extra thread computes the spectrum line.
// demod: 0 none, 1 quadrature, 2 AM magnitude, 3 USB, 4 LSB.
// Returns elapsed seconds for nblocks blocks of `count` samples (<0 on error).
// ---------------------------------------------------------------------------------------------
API double ref_bench_channelizer(double inSR, int nvfo, const double* outSR, const double* bw,
                                 const double* offset, const int* demod, int count, int nblocks,
                                 int nthreads, const complex_t* block,
                                 int fftN, int fftNz, const float* window, int fftFrames) {
    if (nthreads < 1) nthreads = 1;
    struct Chan {
        dsp::channel::RxVFO* vfo = nullptr; dsp::demod::Quadrature* quad = nullptr; SsbObj* ssb = nullptr;
        int demod = 0;
    };
    std::vector<Chan> ch(nvfo);
    for (int v = 0; v < nvfo; v++) {
        ch[v].vfo = new dsp::channel::RxVFO(NULL, inSR, outSR[v], bw[v], offset[v]);
        ch[v].demod = demod[v];
        if (demod[v] == 1) { ch[v].quad = (dsp::demod::Quadrature*)ref_quadrature_create(bw[v] / 2.0, outSR[v]); }
        if (demod[v] == 3 || demod[v] == 4) { ch[v].ssb = (SsbObj*)ref_ssb_create(demod[v] == 3 ? 0 : 1, bw[v], outSR[v]); }
    }
    std::atomic<int> next{0};
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        th.emplace_back([&]() {
            std::vector<complex_t> out(count + 16);
            std::vector<float> dem(count + 16);
            for (;;) {
                int v = next.fetch_add(1);
                if (v >= nvfo) break;
                for (int b = 0; b < nblocks; b++) {
                    int n = ch[v].vfo->process(count, block, out.data());
                    if (ch[v].demod == 1) ch[v].quad->process(n, out.data(), dem.data());
                    else if (ch[v].demod == 2) ref_am_magnitude(n, out.data(), dem.data());
                    else if (ch[v].ssb) ref_ssb_process(ch[v].ssb, n, out.data(), dem.data());
                }
            }
        });
    }
    std::thread fftThread;
    if (fftN > 0 && fftFrames > 0) {
        fftThread = std::thread([&]() {
            std::vector<float> row(fftN);
            for (int f = 0; f < fftFrames; f++) {
                ref_spectrum(fftN, fftNz, block, window, row.data(), nullptr, nullptr);
            }
        });
    }
    for (auto& t : th) t.join();
    if (fftThread.joinable()) fftThread.join();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (auto& c : ch) {
        delete c.vfo;
        if (c.quad) delete c.quad;
        if (c.ssb) delete c.ssb;
    }
    return dt;
}

// fft_scaler (gui/widgets/fft_scaler.h:21-65): the waterfall's zoom / max-decimation of one raw row, as
// WaterFall::pushFFT applies it (gui/widgets/waterfall.cpp:900-904). data must hold fftSize+1 floats.
API void ref_fft_zoom(double viewOffset, double viewBandwidth, double wholeBandwidth, int fftSize, int outSize,
                      const float* data, float* out) {
    fft_scaler sc(viewOffset, viewBandwidth, wholeBandwidth, (size_t)fftSize, (size_t)outSize);
    sc.doZoom(data, out);
}

// ---------------------------------------------------------------------------------------------
// SDR++ server wire packets
// ---------------------------------------------------------------------------------------------
// dsp::compression::SampleStreamDecompressor::process (dsp/compression/sample_stream_decompressor.h:13-36)
API int ref_pcm_decompress(int nbytes, const uint8_t* packet, complex_t* out) {
    dsp::compression::SampleStreamDecompressor d;
    return d.process(nbytes, packet, out);
}
// dsp::compression::SampleStreamCompressor::process (dsp/compression/sample_stream_compressor.h:26-60)
API int ref_pcm_compress(int count, int pcmType, const complex_t* in, uint8_t* packet) {
    return dsp::compression::SampleStreamCompressor::process(count, (dsp::compression::PCMType)pcmType, in, packet);
}

// ---------------------------------------------------------------------------------------------
// radio IF chain blocks (decoder_modules/radio/src/radio_module.h:73-78)
// ---------------------------------------------------------------------------------------------
// dsp::noise_reduction::NoiseBlanker (dsp/noise_reduction/noise_blanker.h)
API void* ref_nb_create(double rate, double level) { return new dsp::noise_reduction::NoiseBlanker(NULL, rate, level); }
API int ref_nb_process(void* h, int count, complex_t* in, complex_t* out) { return ((dsp::noise_reduction::NoiseBlanker*)h)->process(count, in, out); }
API void ref_nb_destroy(void* h) { delete (dsp::noise_reduction::NoiseBlanker*)h; }
// dsp::noise_reduction::Squelch (dsp/noise_reduction/squelch.h). Its block counter is a function-local static: use
// one instance at a time.
API void* ref_squelch_create(double level) { auto* s = new dsp::noise_reduction::Squelch(); s->init(NULL, level); return s; }
API int ref_squelch_process(void* h, int count, complex_t* in, complex_t* out) { return ((dsp::noise_reduction::Squelch*)h)->process(count, in, out); }
API void ref_squelch_destroy(void* h) { delete (dsp::noise_reduction::Squelch*)h; }

// dsp::noise_reduction::FMIF (dsp/noise_reduction/fm_if.h) over oracle/shim/fftw3.h
API void* ref_fmif_create(int bins) { return new dsp::noise_reduction::FMIF(NULL, bins); }
API int ref_fmif_process(void* h, int count, complex_t* in, complex_t* out) { return ((dsp::noise_reduction::FMIF*)h)->process(count, in, out); }
API void ref_fmif_destroy(void* h) { delete (dsp::noise_reduction::FMIF*)h; }

// ---------------------------------------------------------------------------------------------
// the complete demodulators (front end + post-detector stages), float output
// ---------------------------------------------------------------------------------------------
// dsp::demod::FM<float> (dsp/demod/fm.h): Quadrature + optional low-pass FIR
API void* ref_fm_create(double samplerate, double bandwidth, int lowPass) {
    auto* d = new dsp::demod::FM<float>();
    d->init(NULL, samplerate, bandwidth, lowPass != 0, false);
    d->reset();
    return d;
}
API int ref_fm_process(void* h, int count, complex_t* in, float* out) { return ((dsp::demod::FM<float>*)h)->process(count, in, out); }
API void ref_fm_destroy(void* h) { delete (dsp::demod::FM<float>*)h; }

// dsp::demod::AM<float> (dsp/demod/am.h): [carrier AGC] -> magnitude -> DC block -> [audio AGC] -> low-pass FIR
API void* ref_am_create(int agcMode, double bandwidth, double agcAttack, double agcDecay, double dcBlockRate, double samplerate, float agcGain) {
    auto* d = new dsp::demod::AM<float>();
    d->init(NULL, (dsp::demod::AM<float>::AGCMode)agcMode, bandwidth, agcAttack, agcDecay, dcBlockRate, samplerate);
    if (agcGain > 0) d->setAGCGain(agcGain);
    return d;
}
API int ref_am_process(void* h, int count, complex_t* in, float* out) { return ((dsp::demod::AM<float>*)h)->process(count, in, out); }
API void ref_am_destroy(void* h) { delete (dsp::demod::AM<float>*)h; }

// dsp::demod::SSB<float> (dsp/demod/ssb.h): xlate -> real -> AGC. mode 0 USB, 1 LSB, 2 DSB
API void* ref_ssbfull_create(int mode, double bandwidth, double samplerate, int agcEnabled, double agcAttack, double agcDecay) {
    auto* d = new dsp::demod::SSB<float>();
    d->init(NULL, (dsp::demod::SSB<float>::Mode)mode, bandwidth, samplerate, agcEnabled != 0, agcAttack, agcDecay);
    return d;
}
API int ref_ssbfull_process(void* h, int count, const complex_t* in, float* out) { return ((dsp::demod::SSB<float>*)h)->process(count, in, out); }
API void ref_ssbfull_destroy(void* h) { delete (dsp::demod::SSB<float>*)h; }

// dsp::demod::BroadcastFM (dsp/demod/broadcast_fm.h): quadrature demod -> [19 kHz pilot band-pass -> PLL -> L-R down-conversion
// -> L/R matrix] -> [15 kHz low-pass] -> interleaved stereo. out: 2*count floats (l, r, l, r ...). RDS output off.
struct WfmPeek : dsp::demod::BroadcastFM {
    using Base = dsp::demod::BroadcastFM;
    using Base::pilotFirTaps; using Base::audioFirTaps;
};
API void* ref_wfm_create(double deviation, double samplerate, int stereo, int lowPass) {
    auto* d = new dsp::demod::BroadcastFM();
    d->init(NULL, deviation, samplerate, stereo != 0, lowPass != 0, false);
    d->reset();
    return d;
}
API int ref_wfm_process(void* h, int count, complex_t* in, float* out) {
    int rdsCount = 0;
    return ((dsp::demod::BroadcastFM*)h)->process(count, in, (dsp::stereo_t*)out, rdsCount, NULL);
}
// pilot band-pass taps (complex, interleaved) and audio low-pass taps; returns the two tap counts through n[0], n[1]
API void ref_wfm_taps(void* h, int* n, float* pilot, int pcap, float* audio, int acap) {
    WfmPeek* w = static_cast<WfmPeek*>((dsp::demod::BroadcastFM*)h);
    n[0] = (int)w->pilotFirTaps.size; n[1] = (int)w->audioFirTaps.size;
    if (pilot) memcpy(pilot, w->pilotFirTaps.taps, sizeof(complex_t) * std::min<int>(n[0], pcap));
    if (audio) memcpy(audio, w->audioFirTaps.taps, sizeof(float) * std::min<int>(n[1], acap));
}
API void ref_wfm_destroy(void* h) { delete (dsp::demod::BroadcastFM*)h; }
// The same decoder with its RDS side output on (_rdsOut, broadcast_fm.h:168-175,188-198): (mpx, 0) translated by -57 kHz
// (the decoder's own FrequencyXlator) and resampled to 5 kS/s (its own RationalResampler). rds receives the samples of this
// call (at most count of them), *rdsCount their number.
API void* ref_wfm_create_rds(double deviation, double samplerate, int stereo, int lowPass) {
    auto* d = new dsp::demod::BroadcastFM();
    d->init(NULL, deviation, samplerate, stereo != 0, lowPass != 0, true);
    d->reset();
    return d;
}
API int ref_wfm_process_rds(void* h, int count, complex_t* in, float* out, complex_t* rds, int* rdsCount) {
    int n = 0;
    const int rc = ((dsp::demod::BroadcastFM*)h)->process(count, in, (dsp::stereo_t*)out, n, rds);
    if (rdsCount) *rdsCount = n;
    return rc;
}

API const char* ref_build_info() {
#ifdef __FAST_MATH__
    return "reference dsp/ headers + generic-VOLK shim; fast-math timing build";
#else
    return "reference dsp/ headers + generic-VOLK shim; IEEE parity build";
#endif
}
