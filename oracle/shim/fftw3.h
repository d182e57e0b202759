// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for <fftw3.h> (FFTW3f is not installed in this image) so that the
// reference's dsp/noise_reduction/fm_if.h compiles unmodified. Only the calls that header makes are provided. A plan
// is a plain unnormalised DFT of the given sign -- the mathematical definition of what fftwf_plan_dft_1d computes --
// evaluated directly with fp64 twiddles and accumulation and rounded to fp32 once per output (FFTW's own fp32 result
// differs from this by its rounding noise, ~1e-7 relative; parity unpinned at that level).
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstddef>

typedef float fftwf_complex[2];
struct oracle_fftwf_plan_s { int n; const fftwf_complex* in; fftwf_complex* out; int sign; };
typedef oracle_fftwf_plan_s* fftwf_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE (1U << 6)

static inline void* fftwf_malloc(size_t n) { void* p = nullptr; return posix_memalign(&p, 64, n ? n : 64) == 0 ? p : nullptr; }
static inline void fftwf_free(void* p) { free(p); }
static inline fftwf_plan fftwf_plan_dft_1d(int n, fftwf_complex* in, fftwf_complex* out, int sign, unsigned) {
    return new oracle_fftwf_plan_s{ n, in, out, sign };
}
static inline void fftwf_destroy_plan(fftwf_plan p) { delete p; }
static inline void fftwf_execute(const fftwf_plan p) {
    const double step = (double)p->sign * 2.0 * 3.14159265358979323846 / (double)p->n;
    for (int b = 0; b < p->n; b++) {
        double re = 0.0, im = 0.0;
        for (int k = 0; k < p->n; k++) {
            const double a = step * (double)((b * k) % p->n);
            const double c = cos(a), s = sin(a);
            const double xr = p->in[k][0], xi = p->in[k][1];
            re += xr * c - xi * s;
            im += xr * s + xi * c;
        }
        p->out[b][0] = (float)re;
        p->out[b][1] = (float)im;
    }
}
