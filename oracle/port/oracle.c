/* TEST INFRASTRUCTURE ONLY (oracle). Never linked into, imported by or called from the product
 * path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker.
 *
 * Plain-C restatement of the reference's algorithm for the SDR++ signal-path hot loop
 * (SURVEY.md section 8a). Every function cites the reference file:line it follows (paths relative
 * to /root/reference/core/src unless noted). Arithmetic that the reference delegates to VOLK is
 * restated with VOLK's published *generic* kernel semantics (sequential fp32 accumulation;
 * rotator renormalised every 512 samples and at the end of a call) -- VOLK itself is an
 * un-vendored, unpinned dependency (core/CMakeLists.txt:144), as is FFTW3f (:143).
 *
 * Parity pinning: the reference has no tests or golden vectors for this path (SURVEY.md section
 * 4), so this port is pinned against the reference's own headers compiled here
 * (oracle/_ref/libsdrpp_ref.so, tests/test_oracle_vs_ref.py: bit-exact) and against golden
 * vectors generated from that library (tests/golden/, tools/make_golden.py).
 *
 * Build: gcc -std=c11 -O2 -ffp-contract=off (IEEE semantics, no FMA contraction, no fast-math;
 * see SURVEY App. C.1 for why the canon has to be stated).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))
#define ORC_PI 3.14159265358979323846 /* DB_M_PI, dsp/math/constants.h:2 */

typedef struct { float re, im; } cf32; /* dsp::complex_t, dsp/types.h:6-91 */

/* ------------------------------------------------------------------------------------------ */
/* A1. Source-side integer -> cf32 conversions (per scalar, I and Q alike)                     */
/* ------------------------------------------------------------------------------------------ */
enum {
    ORC_FMT_CF32 = 0,
    ORC_FMT_U8_RTL = 1,   /* source_modules/rtl_sdr_source/src/main.cpp:526-527, file_source/src/main.cpp:489 */
    ORC_FMT_U8_TCP = 2,   /* source_modules/rtl_tcp_source/src/rtl_tcp_client.cpp:86-87 */
    ORC_FMT_I8 = 3,       /* volk_8i_s32f_convert_32f(.., 128.0f): hackrf_source/src/main.cpp:386 */
    ORC_FMT_I16_FILE = 4, /* source_modules/file_source/src/main.cpp:506 */
    ORC_FMT_I16_VOLK = 5, /* volk_16i_s32f_convert_32f(.., 32768): bladerf main.cpp:587, plutosdr main.cpp:261-265 */
    ORC_FMT_I24_FILE = 6, /* source_modules/file_source/src/main.cpp:521-527 (packed little-endian 24 bit) */
    ORC_FMT_I32_FILE = 7, /* source_modules/file_source/src/main.cpp:538-544 */
    ORC_FMT_F64 = 8,      /* volk_64f_convert_32f, source_modules/file_source/src/main.cpp:470-476 */
};

API int orc_convert(int fmt, const void* in, int nscalars, float* out) {
    int i;
    switch (fmt) {
    case ORC_FMT_CF32:
        memcpy(out, in, sizeof(float) * (size_t)nscalars);
        return 0;
    case ORC_FMT_U8_RTL: {
        const uint8_t* p = (const uint8_t*)in;
        /* int subtract, float add, fp32 divide -- exactly as the source text reads */
        for (i = 0; i < nscalars; i++) out[i] = (p[i] - 128 + 0.5f) / (128.0f - 0.5f);
        return 0;
    }
    case ORC_FMT_U8_TCP: {
        const uint8_t* p = (const uint8_t*)in;
        for (i = 0; i < nscalars; i++) out[i] = (float)(((double)p[i] - 128.0) / 128.0);
        return 0;
    }
    case ORC_FMT_I8: {
        const int8_t* p = (const int8_t*)in;
        for (i = 0; i < nscalars; i++) out[i] = (float)p[i] / 128.0f;
        return 0;
    }
    case ORC_FMT_I16_FILE: {
        const int16_t* p = (const int16_t*)in;
        for (i = 0; i < nscalars; i++) out[i] = (p[i] + 0.5f) / (32768.0f - 0.5f);
        return 0;
    }
    case ORC_FMT_I16_VOLK: {
        const int16_t* p = (const int16_t*)in;
        for (i = 0; i < nscalars; i++) out[i] = (float)p[i] / 32768.0f;
        return 0;
    }
    case ORC_FMT_I24_FILE: {
        const uint8_t* p = (const uint8_t*)in;
        for (i = 0; i < nscalars; i++, p += 3) {
            int32_t i24 = (int32_t)(((uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16)) << 8) >> 8;
            out[i] = (i24 + 0.5f) / (8388608.0f - 0.5f);
        }
        return 0;
    }
    case ORC_FMT_I32_FILE: {
        const int32_t* p = (const int32_t*)in;
        for (i = 0; i < nscalars; i++) out[i] = (float)((p[i] + 0.5) / (2147483648.0 - 0.5));
        return 0;
    }
    case ORC_FMT_F64: {
        const double* p = (const double*)in;
        for (i = 0; i < nscalars; i++) out[i] = (float)p[i];
        return 0;
    }
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------ */
/* 8f rank 3. SDR++ server wire packets: u16 compression type | u16 PCMType | f32 scaler | data */
/* ------------------------------------------------------------------------------------------ */
enum { ORC_PCM_I8 = 0, ORC_PCM_I16 = 1, ORC_PCM_F32 = 2 }; /* dsp/compression/pcm_type.h:4-8 */

/* dsp/compression/sample_stream_decompressor.h:13-36 (generic VOLK: (float)x / scalar) */
API int orc_pcm_decompress(int nbytes, const uint8_t* packet, cf32* out) {
    uint16_t type;
    float scaler, div;
    float* o = (float*)out;
    int n, i;
    memcpy(&type, packet + 2, 2);
    memcpy(&scaler, packet + 4, 4);
    if (type == ORC_PCM_F32) {
        memcpy(out, packet + 8, (size_t)(nbytes - 8));
        return (nbytes - 8) / (int)sizeof(cf32);
    }
    if (type == ORC_PCM_I16) {
        const int16_t* p = (const int16_t*)(packet + 8);
        n = (nbytes - 8) / 4;
        div = 32768.0f / scaler;
        for (i = 0; i < 2 * n; i++) o[i] = (float)p[i] / div;
        return n;
    }
    if (type == ORC_PCM_I8) {
        const int8_t* p = (const int8_t*)(packet + 8);
        n = (nbytes - 8) / 2;
        div = 128.0f / scaler;
        for (i = 0; i < 2 * n; i++) o[i] = (float)p[i] / div;
        return n;
    }
    return 0;
}

/* dsp/compression/sample_stream_compressor.h:26-60: scaler = first strict maximum of the SIGNED scalars
 * (volk_32f_index_max_32u), payload = saturated rintf(x * (128|32768)/scaler) */
API int orc_pcm_compress(int count, int pcm_type, const cf32* in, uint8_t* packet) {
    const float* x = (const float*)in;
    const uint16_t comp = 0, type = (uint16_t)pcm_type;
    float mx, k, r;
    int i, n = 2 * count;
    memcpy(packet, &comp, 2);
    memcpy(packet + 2, &type, 2);
    if (pcm_type == ORC_PCM_F32) {
        mx = 0.0f;
        memcpy(packet + 4, &mx, 4);
        memcpy(packet + 8, in, (size_t)count * sizeof(cf32));
        return 8 + count * (int)sizeof(cf32);
    }
    mx = x[0];
    for (i = 1; i < n; i++) if (x[i] > mx) mx = x[i];
    memcpy(packet + 4, &mx, 4);
    if (pcm_type == ORC_PCM_I8) {
        int8_t* o = (int8_t*)(packet + 8);
        k = 128.0f / mx;
        for (i = 0; i < n; i++) {
            r = x[i] * k;
            o[i] = r > 127.0f ? (int8_t)127 : r < -128.0f ? (int8_t)-128 : (int8_t)rintf(r);
        }
        return 8 + n;
    }
    if (pcm_type == ORC_PCM_I16) {
        int16_t* o = (int16_t*)(packet + 8);
        k = 32768.0f / mx;
        for (i = 0; i < n; i++) {
            r = x[i] * k;
            if (r > 32767.0f) r = 32767.0f; else if (r < -32768.0f) r = -32768.0f;
            o[i] = (int16_t)rintf(r);
        }
        return 8 + 2 * n;
    }
    return count; /* the reference's fall-through for an unknown type */
}

/* ------------------------------------------------------------------------------------------ */
/* A9. Window design: dsp/window/window.h:38-64, cosine.h:7-16, coefficient headers            */
/* ------------------------------------------------------------------------------------------ */
static double orc_cosine(double n, double N, const double* c, int cnt) {
    /* dsp/window/cosine.h:7-16 */
    double win = 0.0, sign = 1.0;
    int i;
    for (i = 0; i < cnt; i++) {
        win += sign * c[i] * cos((double)i * 2.0 * ORC_PI * n / N);
        sign = -sign;
    }
    return win;
}

static const double C_RECT[] = { 1.0 };                                       /* rectangular.h */
static const double C_HAMMING[] = { 0.53836, 0.46164 };                       /* hamming.h:6 */
static const double C_HANN[] = { 0.5, 0.5 };                                  /* hann.h:6 */
static const double C_BLACKMAN[] = { 0.42, 0.5, 0.08 };                       /* blackman.h:6 */
static const double C_NUTTALL[] = { 0.355768, 0.487396, 0.144232, 0.012604 }; /* nuttall.h:6 */
static const double C_BH4[] = { 0.35875, 0.48829, 0.14128, 0.01168 };         /* blackman_harris4.h:6 */
static const double C_BH7[] = { 0.27105140069342, 0.43329793923448, 0.21812299954311, 0.06592544638803,
                                0.01081174209837, 0.00077658482522, 0.00001388721735 }; /* blackman_harris7.h:22-30 */

/* enum order of dsp::window::windowType, window.h:28-36 */
static int orc_window_coefs(int type, const double** c) {
    switch (type) {
    case 0: *c = C_RECT; return 1;
    case 1: *c = C_HAMMING; return 2;
    case 2: *c = C_HANN; return 2;
    case 3: *c = C_BLACKMAN; return 3;
    case 4: *c = C_NUTTALL; return 4;
    case 5: *c = C_BH4; return 4;
    case 6: *c = C_BH7; return 7;
    }
    return 0;
}

/* createWindow (window.h:38-64). buf needs size+1 floats when centered and size is odd. */
API int orc_window(int type, float* buf, int size, int centered) {
    const double* c;
    int cnt = orc_window_coefs(type, &c), i;
    double wscale = 0.0;
    if (!cnt) return -1;
    for (i = 0; i < size; i++) buf[i] = (float)orc_cosine((double)i, (double)size, c, cnt);
    for (i = 0; i < size; i++) wscale += buf[i];        /* double += float */
    wscale = 1.0 / wscale;
    if (!centered) {
        for (i = 0; i < size; i++) buf[i] = (float)((double)buf[i] * wscale);
    } else {
        for (i = 0; i < size; i += 2) {                 /* float *= double: product in double */
            buf[i] = (float)((double)buf[i] * -wscale);
            buf[i + 1] = (float)((double)buf[i + 1] * wscale);
        }
    }
    return 0;
}

/* A2. genReshapeParams, signal_path/iq_frontend.h:56-60 */
API void orc_reshape_params(double sampleRate, int size, double rate, int* skip, int* nz) {
    int interval = (int)round(sampleRate / rate);
    *nz = interval < size ? interval : size;
    *skip = interval - *nz;
}

/* ------------------------------------------------------------------------------------------ */
/* A5. Tap design: taps/low_pass.h:7-11, windowed_sinc.h:9-29, estimate_tap_count.h:4-6        */
/* ------------------------------------------------------------------------------------------ */
API int orc_lowpass_tap_count(double transWidth, double sampleRate) {
    return (int)(3.8 * sampleRate / transWidth);
}

API int orc_lowpass_taps(double cutoff, double transWidth, double sampleRate, float* out, int cap) {
    int count = orc_lowpass_tap_count(transWidth, sampleRate), i;
    double omega = 2.0 * ORC_PI * (cutoff / sampleRate); /* math/hz_to_rads.h:6-8 */
    double half = (double)count / 2.0;
    double corr = 1.0 * omega / ORC_PI;
    if (!out) return count;
    for (i = 0; i < count && i < cap; i++) {
        double t = (double)i - half + 0.5;
        double x = t * omega;
        double sinc = (x == 0.0) ? 1.0 : (sin(x) / x);  /* math/sinc.h:5-7 */
        out[i] = (float)(sinc * orc_cosine(t - half, (double)count, C_NUTTALL, 4) * corr);
    }
    return count;
}

/* ------------------------------------------------------------------------------------------ */
/* Decimation plans: multirate/decim/plans.h:126-140 (data blob, see tools/extract_decim_plans.py) */
/* ------------------------------------------------------------------------------------------ */
typedef struct { uint32_t len, off; } plan_fir;
typedef struct { uint32_t ratio, nstages; struct { uint32_t decim, fir; } st[4]; } plan_ent;
static struct { int loaded; uint32_t nfirs, nplans, pool_len; plan_fir* firs; plan_ent* plans; float* pool; } g_plans;

API int orc_load_plans(const char* path) {
    FILE* f = fopen(path, "rb");
    uint32_t hdr[5];
    if (!f) return -1;
    if (fread(hdr, 4, 5, f) != 5 || hdr[0] != 0x50445053u || hdr[1] != 1) { fclose(f); return -2; }
    free(g_plans.firs); free(g_plans.plans); free(g_plans.pool);
    g_plans.nfirs = hdr[2]; g_plans.nplans = hdr[3]; g_plans.pool_len = hdr[4];
    g_plans.firs = (plan_fir*)malloc(sizeof(plan_fir) * hdr[2]);
    g_plans.plans = (plan_ent*)malloc(sizeof(plan_ent) * hdr[3]);
    g_plans.pool = (float*)malloc(sizeof(float) * hdr[4]);
    if (fread(g_plans.firs, sizeof(plan_fir), hdr[2], f) != hdr[2] ||
        fread(g_plans.plans, sizeof(plan_ent), hdr[3], f) != hdr[3] ||
        fread(g_plans.pool, sizeof(float), hdr[4], f) != hdr[4]) { fclose(f); return -3; }
    fclose(f);
    g_plans.loaded = 1;
    return 0;
}

/* PowerDecimator::reconfigure plan pick: planId = log2(ratio) - 1 (power_decimator.h:97-98) */
API int orc_decim_plan(int ratio, int* decimation, int* tapcount, const float** taps) {
    uint32_t i, s;
    if (!g_plans.loaded) return -1;
    for (i = 0; i < g_plans.nplans; i++) {
        if ((int)g_plans.plans[i].ratio != ratio) continue;
        for (s = 0; s < g_plans.plans[i].nstages; s++) {
            const plan_fir* fr = &g_plans.firs[g_plans.plans[i].st[s].fir];
            decimation[s] = (int)g_plans.plans[i].st[s].decim;
            tapcount[s] = (int)fr->len;
            taps[s] = g_plans.pool + fr->off;
        }
        return (int)g_plans.plans[i].nstages;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* A4/A14. FIR and decimating FIR: filter/fir.h:62-83, filter/decimating_fir.h:45-68           */
/* State: T-1 past inputs (zero at reset) and the integer `offset` (0 at reset).               */
/* ------------------------------------------------------------------------------------------ */
#ifndef ORC_ACC_T
#define ORC_ACC_T float
#endif
typedef struct {
    int ntaps, decim, offset;
    float* taps;
    cf32* buf;      /* history (ntaps-1) followed by the current block */
    int cap;
} orc_fir;

static void dot_cf(cf32* res, const cf32* x, const float* t, int n) {
    /* volk_32fc_32f_dot_prod_32fc generic: sequential fp32, separate re/im sums. ORC_ACC_T = double builds the
     * "f64" flavour (liboracle_port_f64.so): the same chain with exact-product, fp64-accumulated dot products, used to
     * measure the fp32 scatter of an implementation (the reference's own included) against a common truth. */
    ORC_ACC_T re = 0, im = 0;
    int k;
    for (k = 0; k < n; k++) {
        re += (ORC_ACC_T)x[k].re * (ORC_ACC_T)t[k];
        im += (ORC_ACC_T)x[k].im * (ORC_ACC_T)t[k];
    }
    res->re = (float)re; res->im = (float)im;
}

API orc_fir* orc_fir_create(const float* taps, int ntaps, int decim) {
    orc_fir* f = (orc_fir*)calloc(1, sizeof(orc_fir));
    f->ntaps = ntaps; f->decim = decim < 1 ? 1 : decim; f->offset = 0;
    f->taps = (float*)malloc(sizeof(float) * (size_t)ntaps);
    memcpy(f->taps, taps, sizeof(float) * (size_t)ntaps);
    f->cap = 0; f->buf = NULL;
    return f;
}
static void fir_reserve(orc_fir* f, int count) {
    int need = f->ntaps - 1 + count;
    if (need > f->cap) {
        cf32* nb = (cf32*)calloc((size_t)need + 16, sizeof(cf32));
        if (f->buf) { memcpy(nb, f->buf, sizeof(cf32) * (size_t)(f->ntaps - 1)); free(f->buf); }
        f->buf = nb; f->cap = need;
    }
}
API void orc_fir_reset(orc_fir* f) {
    f->offset = 0;
    if (f->buf) memset(f->buf, 0, sizeof(cf32) * (size_t)(f->ntaps - 1));
}
API int orc_fir_offset(const orc_fir* f) { return f->offset; }
API int orc_fir_process(orc_fir* f, int count, const cf32* in, cf32* out) {
    int n = 0;
    fir_reserve(f, count);
    memcpy(f->buf + (f->ntaps - 1), in, sizeof(cf32) * (size_t)count);
    if (f->decim == 1) {
        /* fir.h:68-78: one output per input */
        for (n = 0; n < count; n++) dot_cf(&out[n], &f->buf[n], f->taps, f->ntaps);
    } else {
        /* decimating_fir.h:51-62 */
        for (; f->offset < count; f->offset += f->decim) dot_cf(&out[n++], &f->buf[f->offset], f->taps, f->ntaps);
        f->offset -= count;
    }
    memmove(f->buf, f->buf + count, sizeof(cf32) * (size_t)(f->ntaps - 1));
    return n;
}
API void orc_fir_destroy(orc_fir* f) { if (f) { free(f->taps); free(f->buf); free(f); } }

/* ------------------------------------------------------------------------------------------ */
/* A3. PowerDecimator: multirate/power_decimator.h:51-67,91-107                                */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int ratio, nstages; orc_fir* st[4]; } orc_pdec;

API orc_pdec* orc_powerdecim_create(int ratio) {
    int dec[4], cnt[4], n, i;
    const float* tp[4];
    orc_pdec* p = (orc_pdec*)calloc(1, sizeof(orc_pdec));
    p->ratio = ratio;
    if (ratio > 1) {
        n = orc_decim_plan(ratio, dec, cnt, tp);
        if (n <= 0) { free(p); return NULL; }
        p->nstages = n;
        for (i = 0; i < n; i++) p->st[i] = orc_fir_create(tp[i], cnt[i], dec[i]);
    }
    return p;
}
API int orc_powerdecim_process(orc_pdec* p, int count, const cf32* in, cf32* out) {
    const cf32* data = in;
    int i;
    if (p->ratio == 1) { memmove(out, in, sizeof(cf32) * (size_t)count); return count; }
    for (i = 0; i < p->nstages; i++) { count = orc_fir_process(p->st[i], count, data, out); data = out; }
    return count;
}
API int orc_powerdecim_offsets(const orc_pdec* p, int* offsets) {
    int i;
    for (i = 0; i < p->nstages; i++) offsets[i] = p->st[i]->offset;
    return p->nstages;
}
API void orc_powerdecim_reset(orc_pdec* p) { int i; for (i = 0; i < p->nstages; i++) orc_fir_reset(p->st[i]); }
API void orc_powerdecim_destroy(orc_pdec* p) { int i; if (!p) return; for (i = 0; i < p->nstages; i++) orc_fir_destroy(p->st[i]); free(p); }

/* ------------------------------------------------------------------------------------------ */
/* A13. Polyphase resampler: multirate/polyphase_bank.h:15-48, polyphase_resampler.h:69-99     */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    int interp, decim, tpp, phase, offset;
    float* bank;    /* [interp][tpp] */
    cf32* buf; int cap;
} orc_poly;

API orc_poly* orc_polyphase_create(int interp, int decim, const float* taps, int ntaps) {
    orc_poly* p = (orc_poly*)calloc(1, sizeof(orc_poly));
    int i, tot;
    p->interp = interp; p->decim = decim;
    p->tpp = (ntaps + interp - 1) / interp;                       /* polyphase_bank.h:24 */
    p->bank = (float*)calloc((size_t)interp * (size_t)p->tpp, sizeof(float));
    tot = interp * p->tpp;
    for (i = 0; i < tot; i++)                                     /* polyphase_bank.h:31-34 */
        p->bank[((interp - 1) - (i % interp)) * p->tpp + (i / interp)] = (i < ntaps) ? taps[i] : 0.0f;
    return p;
}
API int orc_polyphase_tpp(const orc_poly* p) { return p->tpp; }
API void orc_polyphase_state(const orc_poly* p, int* phase, int* offset) { *phase = p->phase; *offset = p->offset; }
API void orc_polyphase_reset(orc_poly* p) {
    p->phase = 0; p->offset = 0;
    if (p->buf) memset(p->buf, 0, sizeof(cf32) * (size_t)(p->tpp - 1));
}
API int orc_polyphase_process(orc_poly* p, int count, const cf32* in, cf32* out) {
    int n = 0, need = p->tpp - 1 + count;
    if (need > p->cap) {
        cf32* nb = (cf32*)calloc((size_t)need + 16, sizeof(cf32));
        if (p->buf) { memcpy(nb, p->buf, sizeof(cf32) * (size_t)(p->tpp - 1)); free(p->buf); }
        p->buf = nb; p->cap = need;
    }
    memcpy(p->buf + (p->tpp - 1), in, sizeof(cf32) * (size_t)count);
    while (p->offset < count) {                                   /* polyphase_resampler.h:75-93 */
        dot_cf(&out[n++], &p->buf[p->offset], p->bank + (size_t)p->phase * (size_t)p->tpp, p->tpp);
        p->phase += p->decim;
        p->offset += p->phase / p->interp;
        p->phase = p->phase % p->interp;
    }
    p->offset -= count;
    memmove(p->buf, p->buf + count, sizeof(cf32) * (size_t)(p->tpp - 1));
    return n;
}
API void orc_polyphase_destroy(orc_poly* p) { if (p) { free(p->bank); free(p->buf); free(p); } }

/* ------------------------------------------------------------------------------------------ */
/* A12. RationalResampler plan + process: multirate/rational_resampler.h:121-167,83-97         */
/* ------------------------------------------------------------------------------------------ */
static int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a < 0 ? -a : a; }

typedef struct {
    int mode;        /* 0 BOTH, 1 DECIM_ONLY, 2 RESAMP_ONLY, 3 NONE (enum order, :114-119) */
    int predec, interp, decim, ntaps;
    orc_pdec* pd; orc_poly* pp;
} orc_resamp;

/* info[0]=mode [1]=predec [2]=interp [3]=decim [4]=ntaps [5]=tpp; taps (optional) are interp-scaled */
API int orc_resampler_plan(double inSR, double outSR, int* info, float* taps, int cap) {
    const int maxRatio = 1 << 13;                                       /* getMaxRatio(), power_decimator.h:28-30 */
    int predecPower = (int)floor(log2(inSR / outSR));
    int predecRatio, useDecim, IntSR, OutSR, g, interp, decim, ntaps, i;
    double intSR = inSR, tapSR, tapBW, tapTW;
    if (predecPower > maxRatio) predecPower = maxRatio;                 /* :123 clamps the exponent against 8192 */
    predecRatio = (predecPower >= 0 && predecPower < 31) ? (1 << predecPower) : ((predecPower < 0) ? 0 : maxRatio);
    if (predecRatio > maxRatio) predecRatio = maxRatio;                 /* :124 */
    useDecim = (inSR > outSR && predecPower > 0);
    if (useDecim) intSR = inSR / (double)predecRatio;
    IntSR = (int)round(intSR); OutSR = (int)round(outSR);
    g = gcd_i(IntSR, OutSR);
    interp = OutSR / g; decim = IntSR / g;
    info[1] = useDecim ? predecRatio : 1;
    if (interp == decim) {
        info[0] = useDecim ? 1 : 3; info[2] = 1; info[3] = 1; info[4] = 0; info[5] = 0;
        return 0;
    }
    tapSR = intSR * (double)interp;
    tapBW = (inSR < outSR ? inSR : outSR) / 2.0;
    tapTW = tapBW * 0.1;
    ntaps = orc_lowpass_taps(tapBW, tapTW, tapSR, taps, taps ? cap : 0);
    if (taps) for (i = 0; i < ntaps && i < cap; i++) taps[i] *= (float)interp;   /* :160 */
    info[0] = useDecim ? 0 : 2; info[2] = interp; info[3] = decim; info[4] = ntaps;
    info[5] = (ntaps + interp - 1) / interp;
    return 0;
}

API orc_resamp* orc_resampler_create(double inSR, double outSR) {
    int info[6];
    orc_resamp* r = (orc_resamp*)calloc(1, sizeof(orc_resamp));
    orc_resampler_plan(inSR, outSR, info, NULL, 0);
    r->mode = info[0]; r->predec = info[1]; r->interp = info[2]; r->decim = info[3]; r->ntaps = info[4];
    if (r->mode == 0 || r->mode == 1) r->pd = orc_powerdecim_create(r->predec);
    if (r->mode == 0 || r->mode == 2) {
        float* t = (float*)malloc(sizeof(float) * (size_t)r->ntaps);
        orc_resampler_plan(inSR, outSR, info, t, r->ntaps);
        r->pp = orc_polyphase_create(r->interp, r->decim, t, r->ntaps);
        free(t);
    }
    return r;
}
API int orc_resampler_process(orc_resamp* r, int count, const cf32* in, cf32* out) {
    switch (r->mode) {
    case 0: count = orc_powerdecim_process(r->pd, count, in, out); return orc_polyphase_process(r->pp, count, out, out);
    case 1: return orc_powerdecim_process(r->pd, count, in, out);
    case 2: return orc_polyphase_process(r->pp, count, in, out);
    default: memmove(out, in, sizeof(cf32) * (size_t)count); return count;
    }
}
API void orc_resampler_reset(orc_resamp* r) { if (r->pd) orc_powerdecim_reset(r->pd); if (r->pp) orc_polyphase_reset(r->pp); }
API void orc_resampler_destroy(orc_resamp* r) { if (!r) return; orc_powerdecim_destroy(r->pd); orc_polyphase_destroy(r->pp); free(r); }

/* ------------------------------------------------------------------------------------------ */
/* A11. FrequencyXlator: channel/frequency_xlator.h:15-23,43-50 on VOLK's generic rotator2     */
/* ------------------------------------------------------------------------------------------ */
/* ideal = 1: the "ideal NCO" flavour: same fp32-quantised increment, but the phase is the closed form
 * n * arg(delta) kept in extended precision instead of the fp32 recurrence; the sample multiply stays fp32. It
 * separates an implementation's own error from the reference rotator's random walk: |ref_f32 - ideal| is the
 * reference's walk, |gpu - ideal| the GPU's error. */
typedef struct { cf32 phase, delta; int ideal; long double turns, phi; } orc_xlat;

static cf32 cmul(cf32 a, cf32 b) { cf32 r; r.re = a.re * b.re - a.im * b.im; r.im = a.re * b.im + a.im * b.re; return r; }

API void orc_xlator_set_offset(orc_xlat* x, double offsetHz, double sampleRate) {
    double w = 2.0 * ORC_PI * (offsetHz / sampleRate);     /* math/hz_to_rads.h:6-8 */
    x->delta.re = (float)cos(w); x->delta.im = (float)sin(w);
    x->turns = (long double)atan2((double)x->delta.im, (double)x->delta.re) / (2.0L * 3.14159265358979323846264338327950288L);
}
API orc_xlat* orc_xlator_create(double offsetHz, double sampleRate) {
    orc_xlat* x = (orc_xlat*)calloc(1, sizeof(orc_xlat));
    x->phase.re = 1.0f; x->phase.im = 0.0f;
    orc_xlator_set_offset(x, offsetHz, sampleRate);
    return x;
}
API orc_xlat* orc_xlator_create_ideal(double offsetHz, double sampleRate) {
    orc_xlat* x = orc_xlator_create(offsetHz, sampleRate);
    x->ideal = 1;
    return x;
}
API void orc_xlator_reset(orc_xlat* x) { x->phase.re = 1.0f; x->phase.im = 0.0f; x->phi = 0.0L; }
API void orc_xlator_state(const orc_xlat* x, float* phase, float* delta) {
    phase[0] = x->phase.re; phase[1] = x->phase.im; delta[0] = x->delta.re; delta[1] = x->delta.im;
}
API int orc_xlator_process(orc_xlat* x, int count, const cf32* in, cf32* out) {
    int i = 0, j, nfull = count / 512, tail = count % 512;
    if (x->ideal) {
        for (i = 0; i < count; i++) {
            const double a = 2.0 * ORC_PI * (double)x->phi;
            cf32 p; p.re = (float)cos(a); p.im = (float)sin(a);
            out[i] = cmul(in[i], p);
            x->phi += x->turns;
            x->phi -= floorl(x->phi + 0.5L);     /* keep |phi| <= 0.5 turns: no precision loss over long runs */
        }
        return count;
    }
    for (j = 0; j < nfull; j++) {
        int k;
        for (k = 0; k < 512; k++, i++) { out[i] = cmul(in[i], x->phase); x->phase = cmul(x->phase, x->delta); }
        { float h = hypotf(x->phase.re, x->phase.im); x->phase.re /= h; x->phase.im /= h; }
    }
    for (j = 0; j < tail; j++, i++) { out[i] = cmul(in[i], x->phase); x->phase = cmul(x->phase, x->delta); }
    if (tail) { float h = hypotf(x->phase.re, x->phase.im); x->phase.re /= h; x->phase.im /= h; }
    return count;
}
API void orc_xlator_destroy(orc_xlat* x) { free(x); }

/* ------------------------------------------------------------------------------------------ */
/* A15. RxVFO: channel/rx_vfo.h:19-33,89-100,117-121                                           */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double inSR, outSR, bw, offset;
    orc_xlat* x; orc_resamp* r; orc_fir* f; int filterNeeded;
} orc_vfo;

static void vfo_make_filter(orc_vfo* v) {
    double fw = v->bw / 2.0;                                   /* generateTaps, rx_vfo.h:117-121 */
    int n = orc_lowpass_taps(fw, fw * 0.1, v->outSR, NULL, 0);
    float* t = (float*)malloc(sizeof(float) * (size_t)n);
    orc_lowpass_taps(fw, fw * 0.1, v->outSR, t, n);
    orc_fir_destroy(v->f);
    v->f = orc_fir_create(t, n, 1);
    fir_reserve(v->f, 1);
    free(t);
}
API orc_vfo* orc_rxvfo_create(double inSR, double outSR, double bw, double offset) {
    orc_vfo* v = (orc_vfo*)calloc(1, sizeof(orc_vfo));
    v->inSR = inSR; v->outSR = outSR; v->bw = bw; v->offset = offset;
    v->filterNeeded = (bw != outSR);
    v->x = orc_xlator_create(-offset, inSR);
    v->r = orc_resampler_create(inSR, outSR);
    vfo_make_filter(v);
    return v;
}
API orc_vfo* orc_rxvfo_create_ideal(double inSR, double outSR, double bw, double offset) {
    orc_vfo* v = orc_rxvfo_create(inSR, outSR, bw, offset);
    v->x->ideal = 1;
    return v;
}
API int orc_rxvfo_process(orc_vfo* v, int count, const cf32* in, cf32* out) {
    orc_xlator_process(v->x, count, in, out);
    count = orc_resampler_process(v->r, count, out, out);
    if (v->filterNeeded) orc_fir_process(v->f, count, out, out);
    return count;
}
API void orc_rxvfo_set_offset(orc_vfo* v, double offset) { v->offset = offset; orc_xlator_set_offset(v->x, -offset, v->inSR); }
API void orc_rxvfo_reset(orc_vfo* v) { orc_xlator_reset(v->x); orc_resampler_reset(v->r); orc_fir_reset(v->f); }
API int orc_rxvfo_info(const orc_vfo* v, int* info) {
    info[0] = v->r->mode; info[1] = v->r->predec; info[2] = v->r->interp; info[3] = v->r->decim; info[4] = v->r->ntaps;
    info[5] = v->r->pp ? v->r->pp->tpp : 0; info[6] = v->filterNeeded ? v->f->ntaps : 0;
    return 0;
}
API void orc_rxvfo_destroy(orc_vfo* v) { if (!v) return; orc_xlator_destroy(v->x); orc_resampler_destroy(v->r); orc_fir_destroy(v->f); free(v); }

/* ------------------------------------------------------------------------------------------ */
/* A5 (DC blocker), A6 (conjugate)                                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float rate; cf32 off; } orc_dcb;
API orc_dcb* orc_dcblock_create(double rate) { orc_dcb* d = (orc_dcb*)calloc(1, sizeof(orc_dcb)); d->rate = (float)rate; return d; }
API int orc_dcblock_process(orc_dcb* d, int count, const cf32* in, cf32* out) {
    int i;                                                     /* correction/dc_blocker.h:54-60 */
    for (i = 0; i < count; i++) {
        cf32 o; o.re = in[i].re - d->off.re; o.im = in[i].im - d->off.im;
        out[i] = o;
        d->off.re += o.re * d->rate; d->off.im += o.im * d->rate;
    }
    return count;
}
API void orc_dcblock_destroy(orc_dcb* d) { free(d); }
API int orc_conjugate(int count, const cf32* in, cf32* out) {
    int i;                                                     /* math/conjugate.h:12-15 */
    for (i = 0; i < count; i++) { out[i].re = in[i].re; out[i].im = -in[i].im; }
    return count;
}

/* ------------------------------------------------------------------------------------------ */
/* 8f rank 4. Radio IF chain blocks (decoder_modules/radio/src/radio_module.h:73-78)           */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float rate, invRate, level, amp; } orc_nb;
API orc_nb* orc_nb_create(double rate, double level) {        /* noise_reduction/noise_blanker.h:12-17,77 */
    orc_nb* b = (orc_nb*)calloc(1, sizeof(orc_nb));
    b->rate = (float)rate; b->invRate = 1.0f - b->rate; b->level = (float)level; b->amp = 1.0f;
    return b;
}
API int orc_nb_process(orc_nb* b, int count, const cf32* in, cf32* out) {
    int i;                                                     /* noise_reduction/noise_blanker.h:39-59 */
    for (i = 0; i < count; i++) {
        float inAmp = sqrtf((in[i].re * in[i].re) + (in[i].im * in[i].im)); /* complex_t::amplitude, types.h:79-81 */
        float gain = 1.0f;
        if (inAmp != 0.0f) {
            float excess;
            b->amp = (b->amp * b->invRate) + (inAmp * b->rate);
            excess = inAmp / b->amp;
            if (excess > b->level) gain = 1.0f / excess;
        }
        out[i].re = in[i].re * gain; out[i].im = in[i].im * gain;
    }
    return count;
}
API void orc_nb_destroy(orc_nb* b) { free(b); }

typedef struct { float level; int mute, cnt; } orc_squelch;  /* cnt: a function-local static in the reference (squelch.h:42) */
API orc_squelch* orc_squelch_create(double level) {
    orc_squelch* s = (orc_squelch*)calloc(1, sizeof(orc_squelch));
    s->level = (float)level;
    return s;
}
API int orc_squelch_process(orc_squelch* s, int count, const cf32* in, cf32* out) {
    float sum = 0.0f, level;                                   /* noise_reduction/squelch.h:34-64 */
    int i;
    for (i = 0; i < count; i++) sum += sqrtf(in[i].re * in[i].re + in[i].im * in[i].im);
    sum /= (float)count;
    level = 20.0f * log10f(sum);
    if (s->mute) {
        if (level < s->level || s->cnt <= 0) s->cnt = 10;
        else if (--s->cnt == 0) s->mute = 0;
    } else if (level < (s->level - 1.0f)) {
        s->cnt = 0;
        s->mute = 1;
    }
    if (!s->mute) memmove(out, in, (size_t)count * sizeof(cf32));
    else memset(out, 0, (size_t)count * sizeof(cf32));
    return count;
}
API void orc_squelch_destroy(orc_squelch* s) { free(s); }

/* FM IF noise reduction (noise_reduction/fm_if.h:45-74): per output sample, window the last `bins` samples with a
 * Nuttall window, forward DFT, keep the strongest bin only, backward DFT, take element bins/2. The DFTs are the
 * mathematical definition (FFTW3f is absent here): fp64 twiddles and sums, one rounding to fp32 per output. */
static const double C_NUTTALL_W[4] = { 0.355768, 0.487396, 0.144232, 0.012604 }; /* window/nuttall.h:6 */
typedef struct { int bins; cf32* hist; float* win; cf32* buf; int cap; } orc_fmif;
API orc_fmif* orc_fmif_create(int bins) {
    orc_fmif* f = (orc_fmif*)calloc(1, sizeof(orc_fmif));
    int i;
    f->bins = bins;
    f->hist = (cf32*)calloc((size_t)bins, sizeof(cf32));
    f->win = (float*)calloc((size_t)bins, sizeof(float));
    for (i = 0; i < bins; i++) f->win[i] = (float)orc_cosine((double)i, (double)(bins - 1), C_NUTTALL_W, 4); /* fm_if.h:111 */
    return f;
}
API int orc_fmif_process(orc_fmif* f, int count, const cf32* in, cf32* out) {
    const int n = f->bins;
    const double fwd = -2.0 * ORC_PI / (double)n, bwd = 2.0 * ORC_PI / (double)n;
    int i, b, k;
    if (f->cap < count + n) { f->cap = count + n; f->buf = (cf32*)realloc(f->buf, (size_t)f->cap * sizeof(cf32)); }
    memcpy(f->buf, f->hist, (size_t)(n - 1) * sizeof(cf32));
    memcpy(f->buf + (n - 1), in, (size_t)count * sizeof(cf32));
    for (i = 0; i < count; i++) {
        float best_amp = 0.0f; cf32 best = { 0.0f, 0.0f }; int idx = 0;
        double a, c, s;
        for (b = 0; b < n; b++) {
            double re = 0.0, im = 0.0;
            cf32 X; float amp;
            for (k = 0; k < n; k++) {
                const double xr = (double)(f->buf[i + k].re * f->win[k]), xi = (double)(f->buf[i + k].im * f->win[k]);
                a = fwd * (double)((b * k) % n); c = cos(a); s = sin(a);
                re += xr * c - xi * s;
                im += xr * s + xi * c;
            }
            X.re = (float)re; X.im = (float)im;
            amp = sqrtf(X.re * X.re + X.im * X.im);
            if (b == 0 || amp > best_amp) { best_amp = amp; best = X; idx = b; } /* volk_32f_index_max_32u: first strict maximum */
        }
        a = bwd * (double)((idx * (n / 2)) % n); c = cos(a); s = sin(a);
        out[i].re = (float)((double)best.re * c - (double)best.im * s);
        out[i].im = (float)((double)best.re * s + (double)best.im * c);
    }
    memcpy(f->hist, f->buf + count, (size_t)(n - 1) * sizeof(cf32));
    return count;
}
API void orc_fmif_destroy(orc_fmif* f) { if (f) { free(f->hist); free(f->win); free(f->buf); free(f); } }

/* ------------------------------------------------------------------------------------------ */
/* A16-A18. Demodulator front ends                                                             */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float invDev; cf32 din; } orc_quad;
API orc_quad* orc_quadrature_create(double deviation, double sampleRate) {
    orc_quad* q = (orc_quad*)calloc(1, sizeof(orc_quad));     /* demod/quadrature.h:21-28; _din zeroed as by reset() */
    q->invDev = (float)(1.0 / (2.0 * ORC_PI * (deviation / sampleRate)));
    return q;
}
API int orc_quadrature_process(orc_quad* q, int count, const cf32* in, float* out) {
    int i;                                                     /* demod/quadrature.h:41-56, USE_QUAD_FM_DEMOD branch */
    for (i = 0; i < count; i++) {
        cf32 y = in[i], c, d;
        c.re = q->din.re; c.im = -q->din.im;
        d = cmul(y, c);                                        /* types.h:23-25 operator* */
        out[i] = atan2f(d.im, d.re) * q->invDev;
        q->din = y;
    }
    return count;
}
API void orc_quadrature_destroy(orc_quad* q) { free(q); }

API int orc_am_magnitude(int count, const cf32* in, float* out) {
    int i;                                                     /* volk_32fc_magnitude_32f, demod/am.h:122 */
    for (i = 0; i < count; i++) out[i] = sqrtf(in[i].re * in[i].re + in[i].im * in[i].im);
    return count;
}

typedef struct { orc_xlat* x; } orc_ssb;
API orc_ssb* orc_ssb_create(int mode, double bandwidth, double sampleRate) {
    orc_ssb* s = (orc_ssb*)calloc(1, sizeof(orc_ssb));        /* demod/ssb.h:119-126: USB +bw/2, LSB -bw/2, DSB 0 */
    double tr = (mode == 0) ? bandwidth / 2.0 : (mode == 1) ? -bandwidth / 2.0 : 0.0;
    s->x = orc_xlator_create(tr, sampleRate);
    return s;
}
API orc_ssb* orc_ssb_create_ideal(int mode, double bandwidth, double sampleRate) {
    orc_ssb* s = orc_ssb_create(mode, bandwidth, sampleRate);
    s->x->ideal = 1;
    return s;
}
API int orc_ssb_process(orc_ssb* s, int count, const cf32* in, float* out) {
    cf32* tmp = (cf32*)malloc(sizeof(cf32) * (size_t)(count + 1));
    int i;                                                     /* demod/ssb.h:90-95; convert/complex_to_real.h:15 */
    orc_xlator_process(s->x, count, in, tmp);
    for (i = 0; i < count; i++) out[i] = tmp[i].re;
    free(tmp);
    return count;
}
API void orc_ssb_destroy(orc_ssb* s) { if (s) { orc_xlator_destroy(s->x); free(s); } }

/* ------------------------------------------------------------------------------------------ */
/* A10. Spectrum line: signal_path/iq_frontend.cpp:230-249 (+ :272-296 zero padding)           */
/* FFTW3f (absent) computes a forward unnormalised DFT; restated as an iterative radix-2 FFT.  */
/* ------------------------------------------------------------------------------------------ */
#define DEF_FFT(NAME, R)                                                                         \
    static void NAME(R* a /* interleaved */, int n) {                                            \
        int i, j, len, k;                                                                        \
        R* w = (R*)malloc(sizeof(R) * (size_t)n);                                                \
        for (i = 1, j = 0; i < n; i++) {                                                         \
            int bit = n >> 1;                                                                    \
            for (; j & bit; bit >>= 1) j ^= bit;                                                 \
            j ^= bit;                                                                            \
            if (i < j) { R tr = a[2*i], ti = a[2*i+1]; a[2*i] = a[2*j]; a[2*i+1] = a[2*j+1]; a[2*j] = tr; a[2*j+1] = ti; } \
        }                                                                                        \
        for (k = 0; k < n / 2; k++) { double ang = -2.0 * ORC_PI * (double)k / (double)n; w[2*k] = (R)cos(ang); w[2*k+1] = (R)sin(ang); } \
        for (len = 2; len <= n; len <<= 1) {                                                     \
            int half = len >> 1, step = n / len;                                                 \
            for (i = 0; i < n; i += len) {                                                       \
                for (k = 0; k < half; k++) {                                                     \
                    R ur = a[2*(i+k)], ui = a[2*(i+k)+1];                                        \
                    R tr = a[2*(i+k+half)], ti = a[2*(i+k+half)+1];                              \
                    R wr = w[2*k*step], wi = w[2*k*step+1];                                      \
                    R vr = tr * wr - ti * wi, vi = tr * wi + ti * wr;                            \
                    a[2*(i+k)] = ur + vr; a[2*(i+k)+1] = ui + vi;                                \
                    a[2*(i+k+half)] = ur - vr; a[2*(i+k+half)+1] = ui - vi;                      \
                }                                                                                \
            }                                                                                    \
        }                                                                                        \
        free(w);                                                                                 \
    }
DEF_FFT(fft_f32, float)
DEF_FFT(fft_f64, double)

/* frame: nz samples, window: nz floats (orc_window(..., centered=1)).
 * row32  (opt): the reference way -- fp32 FFT, then VOLK power spectrum log2(x)*3.0103 with -inf -> -127.
 * X64    (opt): fp64 DFT of the fp32 windowed frame, N interleaved complex doubles.
 * row64  (opt): 10*log10 |X64|^2. */
API int orc_spectrum(int N, int nz, const cf32* frame, const float* window, float* row32, double* X64, double* row64) {
    int i;
    float* u;
    if (N <= 0 || (N & (N - 1)) || nz > N || nz < 0) return -1;
    u = (float*)calloc((size_t)N * 2, sizeof(float));
    for (i = 0; i < nz; i++) { u[2*i] = frame[i].re * window[i]; u[2*i+1] = frame[i].im * window[i]; }  /* :234 */
    if (X64 || row64) {
        double* a = (double*)malloc(sizeof(double) * (size_t)N * 2);
        for (i = 0; i < 2 * N; i++) a[i] = (double)u[i];
        fft_f64(a, N);
        for (i = 0; i < N; i++) {
            if (X64) { X64[2*i] = a[2*i]; X64[2*i+1] = a[2*i+1]; }
            if (row64) row64[i] = 10.0 * log10(a[2*i] * a[2*i] + a[2*i+1] * a[2*i+1]);
        }
        free(a);
    }
    if (row32) {
        fft_f32(u, N);                                                                           /* :237 */
        for (i = 0; i < N; i++) {                                                                /* :244 */
            float re = u[2*i] * 1.0f, im = u[2*i+1] * 1.0f;
            float l = log2f(re * re + im * im);
            if (isinf(l)) l = copysignf(127.0f, l);
            row32[i] = 3.01029995663981209120f * l;
        }
    }
    free(u);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* 8f rank 1. Post-detector stages: loop/agc.h:87-147, correction/dc_blocker.h:54-60 (float),  */
/* filter/fir.h:62-83 (float), demod/fm.h:86-103, demod/am.h:114-146, demod/ssb.h:90-101        */
/* ------------------------------------------------------------------------------------------ */
typedef struct { float setPoint, attack, invAttack, decay, invDecay, maxGain, maxOut, gain, amp; int enabled; } orc_agc;
static void agc_init(orc_agc* a, double attack, double decay) {
    /* AGC::init(NULL, 1.0, attack, decay, 10e6, 10.0, INFINITY) as AM/SSB call it (am.h:32-33, ssb.h:27) */
    float initGain = INFINITY;
    a->setPoint = 1.0f; a->attack = (float)attack; a->invAttack = 1.0f - a->attack;
    a->decay = (float)decay; a->invDecay = 1.0f - a->decay;
    a->maxGain = (float)10e6; a->maxOut = 10.0f;
    a->amp = a->setPoint / initGain;
    a->gain = initGain < a->maxGain ? initGain : a->maxGain;
    a->enabled = 1;
}
static void agc_float(orc_agc* a, int count, const float* in, float* out) {
    int i, j;
    for (i = 0; i < count; i++) {
        if (a->enabled) {
            float inAmp = fabsf(in[i]);
            if (inAmp != 0.0f) {
                a->amp = (inAmp > a->amp) ? ((a->amp * a->invAttack) + (inAmp * a->attack)) : ((a->amp * a->invDecay) + (inAmp * a->decay));
                { float g = a->setPoint / a->amp; a->gain = g < a->maxGain ? g : a->maxGain; }
            } else a->gain = 1.0f;
            if (inAmp * a->gain > a->maxOut) {
                float maxAmp = 0;
                for (j = i; j < count; j++) { float v = fabsf(in[j]); if (v > maxAmp) maxAmp = v; }
                a->amp = maxAmp;
                { float g = a->setPoint / a->amp; a->gain = g < a->maxGain ? g : a->maxGain; }
            }
            out[i] = in[i] * a->gain;
        } else {
            float inAmp = fabsf(in[i]);
            float gainAmp = inAmp * a->gain;
            out[i] = (gainAmp > a->maxOut) ? in[i] * (a->maxOut / inAmp) : in[i] * a->gain;
        }
    }
}
static void agc_complex(orc_agc* a, int count, const cf32* in, cf32* out) {
    int i, j;
    for (i = 0; i < count; i++) { /* carrier AGC is always enabled (am.h:38) */
        float inAmp = sqrtf(in[i].re * in[i].re + in[i].im * in[i].im); /* complex_t::amplitude(), types.h */
        if (inAmp != 0.0f) {
            a->amp = (inAmp > a->amp) ? ((a->amp * a->invAttack) + (inAmp * a->attack)) : ((a->amp * a->invDecay) + (inAmp * a->decay));
            { float g = a->setPoint / a->amp; a->gain = g < a->maxGain ? g : a->maxGain; }
        } else a->gain = 1.0f;
        if (inAmp * a->gain > a->maxOut) {
            float maxAmp = 0;
            for (j = i; j < count; j++) { float v = sqrtf(in[j].re * in[j].re + in[j].im * in[j].im); if (v > maxAmp) maxAmp = v; }
            a->amp = maxAmp;
            { float g = a->setPoint / a->amp; a->gain = g < a->maxGain ? g : a->maxGain; }
        }
        out[i].re = in[i].re * a->gain; out[i].im = in[i].im * a->gain;
    }
}
typedef struct { int ntaps; float* taps; float* buf; int cap; } orc_ffir;
static void ffir_init(orc_ffir* f, const float* taps, int n) {
    f->ntaps = n; f->taps = (float*)malloc(sizeof(float) * (size_t)n); memcpy(f->taps, taps, sizeof(float) * (size_t)n);
    f->buf = NULL; f->cap = 0;
}
static void ffir_process(orc_ffir* f, int count, const float* in, float* out) {
    int need = f->ntaps - 1 + count, i, k;
    if (need > f->cap) {
        float* nb = (float*)calloc((size_t)need + 16, sizeof(float));
        if (f->buf) { memcpy(nb, f->buf, sizeof(float) * (size_t)(f->ntaps - 1)); free(f->buf); }
        f->buf = nb; f->cap = need;
    }
    memcpy(f->buf + (f->ntaps - 1), in, sizeof(float) * (size_t)count);
    for (i = 0; i < count; i++) { /* volk_32f_x2_dot_prod_32f generic: sequential fp32 */
        float acc = 0.0f;
        for (k = 0; k < f->ntaps; k++) acc += f->buf[i + k] * f->taps[k];
        out[i] = acc;
    }
    memmove(f->buf, f->buf + count, sizeof(float) * (size_t)(f->ntaps - 1));
}
static void ffir_free(orc_ffir* f) { free(f->taps); free(f->buf); }

typedef struct {
    int kind;                 /* 1 FM, 2 AM, 3 SSB */
    orc_quad* quad; orc_ssb* ssb;
    int lowpass, agc_mode;    /* AM: 0 OFF 1 CARRIER 2 AUDIO */
    orc_agc audio, carrier; float dc_rate, dc_off;
    orc_ffir lpf; int has_lpf;
    cf32* ctmp; float* ftmp; int cap;
} orc_post;
static void post_reserve(orc_post* p, int n) {
    if (n > p->cap) { free(p->ctmp); free(p->ftmp); p->ctmp = (cf32*)malloc(sizeof(cf32) * (size_t)(n + 16)); p->ftmp = (float*)malloc(sizeof(float) * (size_t)(n + 16)); p->cap = n; }
}
static void post_lpf(orc_post* p, double bandwidth, double samplerate) {
    int n = orc_lowpass_taps(bandwidth / 2.0, (bandwidth / 2.0) * 0.1, samplerate, NULL, 0);
    float* t = (float*)malloc(sizeof(float) * (size_t)n);
    orc_lowpass_taps(bandwidth / 2.0, (bandwidth / 2.0) * 0.1, samplerate, t, n);
    ffir_init(&p->lpf, t, n); p->has_lpf = 1;
    free(t);
}
/* dsp::demod::FM<float>::init(in, samplerate, bandwidth, lowPass, highPass=false), fm.h:25-44,117-145 */
API orc_post* orc_fm_create(double samplerate, double bandwidth, int lowPass) {
    orc_post* p = (orc_post*)calloc(1, sizeof(orc_post));
    p->kind = 1; p->quad = orc_quadrature_create(bandwidth / 2.0, samplerate); p->lowpass = lowPass;
    if (lowPass) post_lpf(p, bandwidth, samplerate);
    return p;
}
/* dsp::demod::AM<float>::init, am.h:27-44 */
API orc_post* orc_am_create(int agcMode, double bandwidth, double agcAttack, double agcDecay, double dcBlockRate, double samplerate, float agcGain) {
    orc_post* p = (orc_post*)calloc(1, sizeof(orc_post));
    p->kind = 2; p->agc_mode = agcMode;
    agc_init(&p->carrier, agcAttack, agcDecay); agc_init(&p->audio, agcAttack, agcDecay);
    p->audio.enabled = (agcMode == 2);
    if (agcGain > 0) p->audio.gain = agcGain;     /* setAGCGain */
    p->dc_rate = (float)dcBlockRate; p->dc_off = 0.0f;
    post_lpf(p, bandwidth, samplerate);
    return p;
}
/* dsp::demod::SSB<float>::init, ssb.h:21-36 */
API orc_post* orc_ssbfull_create(int mode, double bandwidth, double samplerate, int agcEnabled, double agcAttack, double agcDecay) {
    orc_post* p = (orc_post*)calloc(1, sizeof(orc_post));
    p->kind = 3; p->ssb = orc_ssb_create(mode, bandwidth, samplerate);
    agc_init(&p->audio, agcAttack, agcDecay); p->audio.enabled = agcEnabled;
    return p;
}
API int orc_post_process(orc_post* p, int count, const cf32* in, float* out) {
    int i;
    post_reserve(p, count);
    if (p->kind == 1) {
        orc_quadrature_process(p->quad, count, in, out);
        if (p->has_lpf) ffir_process(&p->lpf, count, out, out);
    } else if (p->kind == 2) {
        const cf32* src = in;
        if (p->agc_mode == 1) { agc_complex(&p->carrier, count, in, p->ctmp); src = p->ctmp; }
        orc_am_magnitude(count, src, out);
        for (i = 0; i < count; i++) { float o = out[i] - p->dc_off; out[i] = o; p->dc_off += o * p->dc_rate; } /* dc_blocker.h:54-60 */
        if (p->agc_mode != 1) agc_float(&p->audio, count, out, out);
        ffir_process(&p->lpf, count, out, out);
    } else {
        orc_ssb_process(p->ssb, count, in, out);
        agc_float(&p->audio, count, out, out);
    }
    return count;
}
API void orc_post_destroy(orc_post* p) {
    if (!p) return;
    if (p->quad) orc_quadrature_destroy(p->quad);
    if (p->ssb) orc_ssb_destroy(p->ssb);
    if (p->has_lpf) ffir_free(&p->lpf);
    free(p->ctmp); free(p->ftmp); free(p);
}

/* ------------------------------------------------------------------------------------------ */
/* 8f rank 2. Waterfall zoom / max-decimation: gui/widgets/fft_scaler.h:28-64                   */
/* idx (optional, outSize+1 ints) receives the bin boundaries i0..i_outSize for index parity.   */
/* ------------------------------------------------------------------------------------------ */
API void orc_fft_zoom(double viewOffset, double viewBandwidth, double wholeBandwidth, int fftSize, int outSize,
                      const float* data, float* out, int* idx) {
    const double offsetRatio = viewOffset / (wholeBandwidth / 2.0);
    double width = (viewBandwidth / wholeBandwidth) * fftSize;
    double offset = (((double)fftSize / 2.0) * (offsetRatio + 1)) - (width / 2);
    double factor, f0;
    int i, j;
    if (offset < 0) offset = 0;
    if (width > fftSize - offset) width = fftSize - offset;
    factor = width / outSize;
    f0 = offset;
    if (factor <= 1.0) {
        for (i = 0; i < outSize; i++) {
            int i0 = (int)roundf((float)f0);
            if (idx) idx[i] = i0;
            if (out) out[i] = data[i0];
            f0 = f0 + factor;
        }
        if (idx) idx[outSize] = -1; /* point sampling: no range */
    } else {
        int i0 = (int)roundf((float)f0);
        for (i = 0; i < outSize; i++) {
            double f1 = f0 + factor;
            int i1 = (int)roundf((float)f1);
            if (idx) idx[i] = i0;
            if (out) {
                float m = data[i0];
                for (j = i0 + 1; j < i1; j++) m = (m < data[j]) ? data[j] : m; /* std::max(a,b) = (a<b)?b:a */
                out[i] = m;
            }
            f0 = f1; i0 = i1;
        }
        if (idx) idx[outSize] = i0;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* dsp::demod::BroadcastFM (demod/broadcast_fm.h), the WFM stereo decoder and its RDS side output */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double deviation, samplerate; int stereo, lowPass;
    orc_quad* demod;
    int np, na, delay;
    cf32* ptaps; float* ataps;
    float pll_alpha, pll_beta, pll_phase, pll_freq, pll_init_freq, pll_min, pll_max;
    float* pbuf; float* dbuf; float* lbuf; float* rbuf;   /* [hist | block] buffers, grown on demand */
    int cap;
    int rds;                 /* _rdsOut: xlator.init(NULL, -57000.0, samplerate), rdsResamp.init(NULL, samplerate, 5000.0) (broadcast_fm.h:50-51) */
    orc_xlat* rx; orc_resamp* rr;
} orc_wfm;
static double nuttall_d(double n, double N) {
    static const double c[4] = { 0.355768, 0.487396, 0.144232, 0.012604 };
    return orc_cosine(n, N, c, 4);
}
/* taps::bandPass<complex_t>(bandStart, bandStop, transWidth, sampleRate, oddTapCount): taps/band_pass.h:10-25, windowed_sinc.h:9-29 */
static cf32* bandpass_complex(double bandStart, double bandStop, double transWidth, double sampleRate, int odd, int* count_out) {
    float offsetOmega = (float)(2.0 * ORC_PI * (((bandStart + bandStop) / 2.0) / sampleRate));
    int count = (int)(3.8 * sampleRate / transWidth), i;
    double omega = 2.0 * ORC_PI * (((bandStop - bandStart) / 2.0) / sampleRate), half, corr;
    cf32* taps;
    if (odd && !(count % 2)) count++;
    half = (double)count / 2.0; corr = 1.0 * omega / ORC_PI;
    taps = (cf32*)malloc(sizeof(cf32) * (size_t)count);
    for (i = 0; i < count; i++) {
        double t = (double)i - half + 0.5, x = t * omega, n = t - half;
        cf32 cplx, w, r;
        float wn = (float)nuttall_d(n, (double)count), ph = -offsetOmega * (float)n;
        cplx.re = (float)((x == 0.0) ? 1.0 : (sin(x) / x)); cplx.im = 0.0f;
        w.re = cosf(ph) * wn; w.im = sinf(ph) * wn;
        r.re = (cplx.re * w.re) - (cplx.im * w.im); r.im = (cplx.im * w.re) + (cplx.re * w.im);
        taps[i].re = r.re * (float)corr; taps[i].im = r.im * (float)corr;
    }
    *count_out = count;
    return taps;
}
API orc_wfm* orc_wfm_create(double deviation, double samplerate, int stereo, int lowPass) {
    orc_wfm* w = (orc_wfm*)calloc(1, sizeof(orc_wfm));
    float bw = (float)(25000.0 / samplerate), damp = (float)(sqrt(2.0) / 2.0), den;
    w->deviation = deviation; w->samplerate = samplerate; w->stereo = stereo; w->lowPass = lowPass;
    w->demod = orc_quadrature_create(deviation, samplerate);
    w->ptaps = bandpass_complex(18750.0, 19250.0, 3000.0, samplerate, 1, &w->np);
    w->na = orc_lowpass_taps(15000.0, 4000.0, samplerate, NULL, 0);
    w->ataps = (float*)malloc(sizeof(float) * (size_t)w->na);
    orc_lowpass_taps(15000.0, 4000.0, samplerate, w->ataps, w->na);
    w->delay = ((w->np - 1) / 2) + 1;
    den = (float)(1.0 + 2.0 * damp * bw + bw * bw);          /* PhaseControlLoop<float>::criticallyDamped */
    w->pll_alpha = (4 * damp * bw) / den; w->pll_beta = (4 * bw * bw) / den;
    w->pll_init_freq = (float)(2.0 * ORC_PI * (19000.0 / samplerate));
    w->pll_min = (float)(2.0 * ORC_PI * (18750.0 / samplerate)); w->pll_max = (float)(2.0 * ORC_PI * (19250.0 / samplerate));
    w->pll_phase = 0.0f; w->pll_freq = w->pll_init_freq;
    return w;
}
/* rds = 1: the reference's own rotator; rds = 2: the ideal-NCO flavour of the translation, everything else the same */
API orc_wfm* orc_wfm_create_rds(double deviation, double samplerate, int stereo, int lowPass, int rds) {
    orc_wfm* w = orc_wfm_create(deviation, samplerate, stereo, lowPass);
    w->rds = rds;
    if (rds) {
        w->rx = (rds == 2) ? orc_xlator_create_ideal(-57000.0, samplerate) : orc_xlator_create(-57000.0, samplerate);
        w->rr = orc_resampler_create(samplerate, 5000.0);
    }
    return w;
}
API void orc_wfm_taps(const orc_wfm* w, int* n, float* pilot, int pcap, float* audio, int acap) {
    n[0] = w->np; n[1] = w->na;
    if (pilot) memcpy(pilot, w->ptaps, sizeof(cf32) * (size_t)(w->np < pcap ? w->np : pcap));
    if (audio) memcpy(audio, w->ataps, sizeof(float) * (size_t)(w->na < acap ? w->na : acap));
}
static void wfm_reserve(orc_wfm* w, int n) {
    if (n <= w->cap) return;
    {
        float* nb[4]; int hist[4], i;
        float** old[4];
        old[0] = &w->pbuf; old[1] = &w->dbuf; old[2] = &w->lbuf; old[3] = &w->rbuf;
        hist[0] = w->np - 1; hist[1] = w->delay; hist[2] = w->na - 1; hist[3] = w->na - 1;
        for (i = 0; i < 4; i++) {
            nb[i] = (float*)calloc((size_t)(hist[i] + n + 16), sizeof(float));
            if (*old[i]) { memcpy(nb[i], *old[i], sizeof(float) * (size_t)hist[i]); free(*old[i]); }
            *old[i] = nb[i];
        }
        w->cap = n;
    }
}
static void ffir_inplace(float* buf, int hist, int n, const float* taps, int nt, float* out) {
    int i, k;                                  /* FIR<float,float>::process, fir.h:62-83; volk_32f_x2_dot_prod_32f generic */
    for (i = 0; i < n; i++) { ORC_ACC_T acc = 0; for (k = 0; k < nt; k++) acc += (ORC_ACC_T)buf[i + k] * (ORC_ACC_T)taps[k]; out[i] = (float)acc; }
    memmove(buf, buf + n, sizeof(float) * (size_t)hist);
}
/* BroadcastFM::process, broadcast_fm.h:147-214; out = interleaved stereo_t (l, r) */
static int wfm_process(orc_wfm* w, int count, const cf32* in, float* out, cf32* rds, int* rdsCount) {
    const float PI = 3.1415926535f;            /* FL_M_PI */
    float* mpx = (float*)malloc(sizeof(float) * (size_t)(count + 1));
    int i, k;
    orc_quadrature_process(w->demod, count, in, mpx);
    wfm_reserve(w, count);
    if (rdsCount) *rdsCount = 0;
    if (w->rds && rds) {
        /* rtoc (convert/real_to_complex.h: (x, 0)) -> xlator -> rdsResamp, broadcast_fm.h:168-175 (stereo) / 188-198 (mono): the
         * same values either way -- the stereo branch translates rtoc's buffer in place AFTER the delay line has copied it */
        cf32* c = (cf32*)malloc(sizeof(cf32) * (size_t)(count + 1));
        int n;
        for (i = 0; i < count; i++) { c[i].re = mpx[i]; c[i].im = 0.0f; }
        orc_xlator_process(w->rx, count, c, c);
        n = orc_resampler_process(w->rr, count, c, rds);
        if (rdsCount) *rdsCount = n;
        free(c);
    }
    if (w->stereo) {
        cf32* pil = (cf32*)malloc(sizeof(cf32) * (size_t)(count + 1));
        float* l = w->lbuf + (w->na - 1); float* r = w->rbuf + (w->na - 1);
        memcpy(w->pbuf + (w->np - 1), mpx, sizeof(float) * (size_t)count);
        for (i = 0; i < count; i++) {          /* pilotFir on rtoc's (mpx, 0): volk_32fc_x2_dot_prod_32fc generic */
            ORC_ACC_T re = 0, im = 0;
            for (k = 0; k < w->np; k++) {
                const ORC_ACC_T ar = w->pbuf[i + k], ai = 0.0f, br = w->ptaps[k].re, bi = w->ptaps[k].im;
                re += ar * br - ai * bi; im += ar * bi + ai * br;
            }
            pil[i].re = (float)re; pil[i].im = (float)im;
        }
        memmove(w->pbuf, w->pbuf + count, sizeof(float) * (size_t)(w->np - 1));
        memcpy(w->dbuf + w->delay, mpx, sizeof(float) * (size_t)count);   /* lprDelay / lmrDelay: the same delayed mpx */
        for (i = 0; i < count; i++) {
            cf32 vco, c, m1, m2; float err, d = w->dbuf[i], lmr;
            vco.re = cosf(w->pll_phase); vco.im = sinf(w->pll_phase);                     /* loop/pll.h:66-72 */
            err = atan2f(pil[i].im, pil[i].re) - w->pll_phase;
            if (err > PI) err -= 2.0f * PI; else if (err <= -PI) err += 2.0f * PI;      /* math/normalize_phase.h */
            w->pll_freq += w->pll_beta * err;                                             /* phase_control_loop.h:58-66 */
            if (w->pll_freq > w->pll_max) w->pll_freq = w->pll_max; else if (w->pll_freq < w->pll_min) w->pll_freq = w->pll_min;
            w->pll_phase += w->pll_freq + (w->pll_alpha * err);
            while (w->pll_phase > PI) w->pll_phase -= (PI - (-PI));
            while (w->pll_phase < -PI) w->pll_phase += (PI - (-PI));
            c.re = vco.re; c.im = -vco.im;                                                /* math/conjugate.h */
            m1.re = d * c.re - 0.0f * c.im; m1.im = d * c.im + 0.0f * c.re;               /* volk_32fc_x2_multiply_32fc, twice */
            m2.re = m1.re * c.re - m1.im * c.im;
            lmr = m2.re * 2.0f;
            l[i] = d + lmr; r[i] = d - lmr;
        }
        memmove(w->dbuf, w->dbuf + count, sizeof(float) * (size_t)w->delay);
        if (w->lowPass) {
            float* tl = (float*)malloc(sizeof(float) * (size_t)(count + 1)); float* tr = (float*)malloc(sizeof(float) * (size_t)(count + 1));
            ffir_inplace(w->lbuf, w->na - 1, count, w->ataps, w->na, tl);
            ffir_inplace(w->rbuf, w->na - 1, count, w->ataps, w->na, tr);
            for (i = 0; i < count; i++) { out[2 * i] = tl[i]; out[2 * i + 1] = tr[i]; }
            free(tl); free(tr);
        } else {
            for (i = 0; i < count; i++) { out[2 * i] = l[i]; out[2 * i + 1] = r[i]; }
        }
        free(pil);
    } else {
        if (w->lowPass) {
            float* t = (float*)malloc(sizeof(float) * (size_t)(count + 1));
            memcpy(w->lbuf + (w->na - 1), mpx, sizeof(float) * (size_t)count);
            ffir_inplace(w->lbuf, w->na - 1, count, w->ataps, w->na, t);
            for (i = 0; i < count; i++) { out[2 * i] = t[i]; out[2 * i + 1] = t[i]; }
            free(t);
        } else {
            for (i = 0; i < count; i++) { out[2 * i] = mpx[i]; out[2 * i + 1] = mpx[i]; }
        }
    }
    free(mpx);
    return count;
}
API int orc_wfm_process(orc_wfm* w, int count, const cf32* in, float* out) { return wfm_process(w, count, in, out, NULL, NULL); }
API int orc_wfm_process_rds(orc_wfm* w, int count, const cf32* in, float* out, cf32* rds, int* rdsCount) { return wfm_process(w, count, in, out, rds, rdsCount); }
API void orc_wfm_destroy(orc_wfm* w) {
    if (!w) return;
    orc_xlator_destroy(w->rx); orc_resampler_destroy(w->rr);
    orc_quadrature_destroy(w->demod); free(w->ptaps); free(w->ataps); free(w->pbuf); free(w->dbuf); free(w->lbuf); free(w->rbuf); free(w);
}

/* ------------------------------------------------------------------------------------------ */
/* level / SNR read-out and the waterfall's per-line display state              */
/* ------------------------------------------------------------------------------------------ */
/* WaterFall::calculateVFOSignalInfo, gui/widgets/waterfall.cpp:563-603 (restated: the widget cannot be compiled here) */
static int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
API int orc_vfo_signal_info(const float* fftLine, int rawFFTSize, double centerOffset, double bandwidth, double wholeBandwidth,
                            float* strength, float* snr) {
    double vfoMinSizeFreq = centerOffset - bandwidth;
    double vfoMinFreq = centerOffset - (bandwidth / 2.0);
    double vfoMaxFreq = centerOffset + (bandwidth / 2.0);
    double vfoMaxSizeFreq = centerOffset + bandwidth;
    int vfoMinSideOffset = clampi((int)(((vfoMinSizeFreq / (wholeBandwidth / 2.0)) * (double)(rawFFTSize / 2)) + (rawFFTSize / 2)), 0, rawFFTSize);
    int vfoMinOffset = clampi((int)(((vfoMinFreq / (wholeBandwidth / 2.0)) * (double)(rawFFTSize / 2)) + (rawFFTSize / 2)), 0, rawFFTSize);
    int vfoMaxOffset = clampi((int)(((vfoMaxFreq / (wholeBandwidth / 2.0)) * (double)(rawFFTSize / 2)) + (rawFFTSize / 2)), 0, rawFFTSize);
    int vfoMaxSideOffset = clampi((int)(((vfoMaxSizeFreq / (wholeBandwidth / 2.0)) * (double)(rawFFTSize / 2)) + (rawFFTSize / 2)), 0, rawFFTSize);
    double avg = 0;
    float max = -INFINITY;
    int avgCount = 0, i;
    if (!fftLine) return 0;
    for (i = vfoMinSideOffset; i < vfoMinOffset; i++) { avg += fftLine[i]; avgCount++; }
    for (i = vfoMaxOffset + 1; i < vfoMaxSideOffset; i++) { avg += fftLine[i]; avgCount++; }
    avg /= (double)(avgCount);
    for (i = vfoMinOffset; i <= vfoMaxOffset && i < rawFFTSize; i++) { if (fftLine[i] > max) max = fftLine[i]; }   /* the reference reads fftLine[rawFFTSize] when the VFO touches the upper edge */
    *strength = max;
    *snr = (float)(max - avg);
    return 1;
}
/* WaterFall::pushFFT, waterfall.cpp:918-925 (smoothing: three generic-VOLK calls) and :951-956 (peak hold), on nrows zoomed
 * rows of dataWidth pixels in place; smoothingBuf / latestFFTHold carry from call to call. */
API void orc_fft_display(int dataWidth, int nrows, float* rows, int smoothing, float alpha, float* smoothingBuf, int hold, float holdSpeed,
                         float* latestFFTHold) {
    const float beta = 1.0f - alpha;
    int r, i;
    for (r = 0; r < nrows; r++) {
        float* latestFFT = rows + (size_t)r * (size_t)dataWidth;
        if (smoothing) {
            for (i = 0; i < dataWidth; i++) latestFFT[i] = latestFFT[i] * alpha;            /* volk_32f_s32f_multiply_32f */
            for (i = 0; i < dataWidth; i++) smoothingBuf[i] = smoothingBuf[i] * beta;       /* volk_32f_s32f_multiply_32f */
            for (i = 0; i < dataWidth; i++) smoothingBuf[i] = smoothingBuf[i] + latestFFT[i]; /* volk_32f_x2_add_32f */
            memcpy(latestFFT, smoothingBuf, sizeof(float) * (size_t)dataWidth);
        }
        if (hold) {
            for (i = 1; i < dataWidth; i++) {
                const float d = latestFFTHold[i] - holdSpeed;
                latestFFTHold[i] = (latestFFT[i] < d) ? d : latestFFT[i];                   /* std::max<float>(latestFFT[i], hold - speed) */
            }
        }
    }
}

API const char* orc_build_info(void) { return "oracle port: plain-C restatement, IEEE fp32, generic-VOLK semantics"; }
