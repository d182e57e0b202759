// TEST INFRASTRUCTURE ONLY (oracle). Not part of the product path.
//
// Stand-in for <volk/volk.h>, which the reference includes but does not vendor
// (core/CMakeLists.txt:144: pkg_check_modules(VOLK REQUIRED volk), unpinned).
// Each function restates the published *generic* (scalar C) VOLK kernel so the
// reference's own dsp/ headers compile unmodified into oracle/_ref.
// Accumulation type is selectable so the reference's own fp32 rounding noise can
// be separated from GPU error: -DORACLE_ACC_T=double builds the "ref_f64" flavour.
#pragma once
#include <complex>
#include <cmath>
#include <cstdlib>
#include <cstdint>
#include <cstring>

#ifndef ORACLE_ACC_T
#define ORACLE_ACC_T float
#endif

#define VOLK_VERSION 030100   // selects the rotator2 entry point (frequency_xlator.h:44)

typedef std::complex<float> lv_32fc_t;
#define lv_cmake(r, i) lv_32fc_t((float)(r), (float)(i))
static inline float lv_creal(const lv_32fc_t& x) { return x.real(); }
static inline float lv_cimag(const lv_32fc_t& x) { return x.imag(); }

static inline size_t volk_get_alignment() { return 64; }
static inline void* volk_malloc(size_t size, size_t alignment) {
    void* p = nullptr;
    if (size == 0) size = alignment;
    if (posix_memalign(&p, alignment, size) != 0) return nullptr;
    return p;
}
static inline void volk_free(void* p) { free(p); }

// --- dot products: sequential accumulate, separate re/im accumulators -----------------
static inline void volk_32fc_32f_dot_prod_32fc(lv_32fc_t* result, const lv_32fc_t* input,
                                               const float* taps, unsigned int num_points) {
    const float* a = (const float*)input;
    ORACLE_ACC_T re = 0, im = 0;
    for (unsigned int n = 0; n < num_points; n++) {
        re += (ORACLE_ACC_T)a[2 * n] * (ORACLE_ACC_T)taps[n];
        im += (ORACLE_ACC_T)a[2 * n + 1] * (ORACLE_ACC_T)taps[n];
    }
    *result = lv_32fc_t((float)re, (float)im);
}
static inline void volk_32f_x2_dot_prod_32f(float* result, const float* input, const float* taps,
                                            unsigned int num_points) {
    ORACLE_ACC_T acc = 0;
    for (unsigned int n = 0; n < num_points; n++) acc += (ORACLE_ACC_T)input[n] * (ORACLE_ACC_T)taps[n];
    *result = (float)acc;
}
static inline void volk_32fc_x2_dot_prod_32fc(lv_32fc_t* result, const lv_32fc_t* input,
                                              const lv_32fc_t* taps, unsigned int num_points) {
    ORACLE_ACC_T re = 0, im = 0;
    for (unsigned int n = 0; n < num_points; n++) {
        const ORACLE_ACC_T ar = input[n].real(), ai = input[n].imag();
        const ORACLE_ACC_T br = taps[n].real(), bi = taps[n].imag();
        re += ar * br - ai * bi;
        im += ar * bi + ai * br;
    }
    *result = lv_32fc_t((float)re, (float)im);
}

// --- rotator (generic): per-sample fp32 recurrence, renormalise every 512 samples and at
//     the end of a call with a partial tail ---------------------------------------------
#define ORACLE_ROTATOR_RELOAD 512
static inline lv_32fc_t oracle_cmul(const lv_32fc_t& a, const lv_32fc_t& b) {
    return lv_32fc_t(a.real() * b.real() - a.imag() * b.imag(), a.real() * b.imag() + a.imag() * b.real());
}
static inline void volk_32fc_s32fc_x2_rotator2_32fc(lv_32fc_t* out, const lv_32fc_t* in,
                                                    const lv_32fc_t* phase_inc, lv_32fc_t* phase,
                                                    unsigned int num_points) {
    unsigned int i = 0;
    for (i = 0; i < num_points / ORACLE_ROTATOR_RELOAD; ++i) {
        for (int j = 0; j < ORACLE_ROTATOR_RELOAD; ++j) {
            *out++ = oracle_cmul(*in++, *phase);
            *phase = oracle_cmul(*phase, *phase_inc);
        }
        *phase /= hypotf(phase->real(), phase->imag());
    }
    for (i = 0; i < num_points % ORACLE_ROTATOR_RELOAD; ++i) {
        *out++ = oracle_cmul(*in++, *phase);
        *phase = oracle_cmul(*phase, *phase_inc);
    }
    if (i) { *phase /= hypotf(phase->real(), phase->imag()); }
}
static inline void volk_32fc_s32fc_x2_rotator_32fc(lv_32fc_t* out, const lv_32fc_t* in,
                                                   const lv_32fc_t phase_inc, lv_32fc_t* phase,
                                                   unsigned int num_points) {
    volk_32fc_s32fc_x2_rotator2_32fc(out, in, &phase_inc, phase, num_points);
}

// --- element-wise ---------------------------------------------------------------------------
static inline void volk_32fc_conjugate_32fc(lv_32fc_t* out, const lv_32fc_t* in, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) out[i] = std::conj(in[i]);
}
static inline void volk_32fc_magnitude_32f(float* out, const lv_32fc_t* in, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) {
        const float re = in[i].real(), im = in[i].imag();
        out[i] = sqrtf(re * re + im * im);
    }
}
static inline void volk_32fc_deinterleave_real_32f(float* out, const lv_32fc_t* in, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) out[i] = in[i].real();
}
static inline void volk_32f_x2_interleave_32fc(lv_32fc_t* out, const float* re, const float* im, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) out[i] = lv_32fc_t(re[i], im[i]);
}
static inline void volk_32fc_32f_multiply_32fc(lv_32fc_t* out, const lv_32fc_t* a, const float* b, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) out[i] = lv_32fc_t(a[i].real() * b[i], a[i].imag() * b[i]);
}
static inline void volk_32fc_x2_multiply_32fc(lv_32fc_t* out, const lv_32fc_t* a, const lv_32fc_t* b, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) out[i] = oracle_cmul(a[i], b[i]);
}
static inline void volk_32f_x2_multiply_32f(float* out, const float* a, const float* b, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) out[i] = a[i] * b[i];
}
static inline void volk_32f_x2_add_32f(float* out, const float* a, const float* b, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) out[i] = a[i] + b[i];
}
static inline void volk_32f_x2_subtract_32f(float* out, const float* a, const float* b, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) out[i] = a[i] - b[i];
}
static inline void volk_32f_s32f_multiply_32f(float* out, const float* a, const float s, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) out[i] = a[i] * s;
}

// --- power spectrum: 10*log10(|X|^2 / norm^2) as log2 * (10/log2(10)); -inf clamps to -127 --
static inline float oracle_log2f_non_ieee(float f) {
    const float r = log2f(f);
    return std::isinf(r) ? copysignf(127.0f, r) : r;
}
static inline void volk_32fc_s32f_power_spectrum_32f(float* logPower, const lv_32fc_t* in,
                                                     const float normalizationFactor, unsigned int n) {
    const float inv = 1.0f / normalizationFactor;
    for (unsigned int i = 0; i < n; i++) {
        const float re = in[i].real() * inv, im = in[i].imag() * inv;
        logPower[i] = 3.01029995663981209120f * oracle_log2f_non_ieee(re * re + im * im);
    }
}

// --- integer -> float conversions: (float)x / scale ------------------------------------------
static inline void volk_16i_s32f_convert_32f(float* out, const int16_t* in, const float scale, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) out[i] = (float)in[i] / scale;
}
static inline void volk_8i_s32f_convert_32f(float* out, const int8_t* in, const float scale, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) out[i] = (float)in[i] / scale;
}

// --- float -> integer conversions (generic VOLK: r = x * scale; saturate; rintf) and index of the
// --- maximum (first strict maximum), used by dsp/compression/sample_stream_compressor.h -------
static inline void volk_32f_s32f_convert_8i(int8_t* out, const float* in, const float scale, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) {
        const float r = in[i] * scale;
        out[i] = r > 127.0f ? (int8_t)127 : r < -128.0f ? (int8_t)-128 : (int8_t)rintf(r);
    }
}
static inline void volk_32f_s32f_convert_16i(int16_t* out, const float* in, const float scale, unsigned int n) {
    for (unsigned int i = 0; i < n; i++) {
        float r = in[i] * scale;
        if (r > 32767.0f) r = 32767.0f; else if (r < -32768.0f) r = -32768.0f;
        out[i] = (int16_t)rintf(r);
    }
}
// generic volk_32f_accumulator_s32f: in-order fp32 sum (noise_reduction/squelch.h:37)
static inline void volk_32f_accumulator_s32f(float* result, const float* in, unsigned int n) {
    float acc = 0.0f;
    for (unsigned int i = 0; i < n; i++) acc += in[i];
    *result = acc;
}
static inline void volk_32f_index_max_32u(uint32_t* target, const float* src, uint32_t n) {
    if (n == 0) return;
    float mx = src[0];
    uint32_t idx = 0;
    for (uint32_t i = 1; i < n; i++) if (src[i] > mx) { idx = i; mx = src[i]; }
    *target = idx;
}
