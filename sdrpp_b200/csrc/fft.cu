// Fused spectrum-line kernels: window * frame -> forward DFT -> 10*log10|X|^2 row.
// Replaces IQFrontEnd::handler (signal_path/iq_frontend.cpp:230-249): volk_32fc_32f_multiply_32fc
// + fftwf_execute + volk_32fc_s32f_power_spectrum_32f.
//
// N <= 4096: one kernel, one CTA per group of frames, data stays in registers/shared memory.
// N >= 8192: four-step decomposition N = N1*N2 in two kernels:
//   cols kernel: for each column n2, DFT over n1 of x[n1*N2+n2]*w[...] (window and zero padding
//                fused into the strided load), times W_N^(n2*k1), written to A[k1][n2];
//   rows kernel: for each row k1, DFT over n2 of A[k1][.] -> X[k1+N1*k2] -> dB -> row.
// The intermediate A (8 B/sample) is written and re-read within microseconds and stays in the
// 126 MB L2; HBM sees the 8 B/sample read and the 4 B/sample row write.
#include "fft_core.cuh"
#include "kernels.h"
#include "../../include/sdrpp_cuda.h"

namespace sdrpp {

__device__ __forceinline__ float power_db(float2 x) {
    // volk_32fc_s32f_power_spectrum_32f with norm 1: log2(re^2+im^2) * 10/log2(10), -inf -> -127*3.0103
    float l = __log2f(x.x * x.x + x.y * x.y);
    if (isinf(l)) l = copysignf(127.0f, l);
    return 3.01029995663981209120f * l;
}

// ---------------------------------------------------------------------------------------------
// cols kernel: grid (N2/B, frames), block T*B, thread (t,b) with b fastest (coalesced columns)
// ---------------------------------------------------------------------------------------------
template <class P, int B>
__global__ void __launch_bounds__(P::T* B)
fft_cols_kernel(SpectrumArgs a, int N2, int log2N) {
    extern __shared__ __align__(16) float2 sm[];
    constexpr int E = P::E, T = P::T;
    const int b = threadIdx.x % B, t = threadIdx.x / B;
    const int n2 = blockIdx.x * B + b;
    const int f = blockIdx.y;
    const uint32_t base = a.start + (uint32_t)f * a.frame_stride;
    const float2* __restrict__ in = reinterpret_cast<const float2*>(a.in);

    float2 v[E];
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int n = (t + T * e) * N2 + n2;
        if (n < a.nz) {
            const float2 x = __ldg(in + ((base + (uint32_t)n) & a.ring_mask));
            const float w = __ldg(a.window + n);
            v[e] = make_float2(x.x * w, x.y * w);
        } else {
            v[e] = make_float2(0.0f, 0.0f);
        }
    }
    block_fft<P, true, B>(v, sm, t, b);

    // four-step twiddle W_N^(n2*k1), k1 = t + T*e: base * step^e
    const uint32_t nmask = (1u << log2N) - 1u;
    float sn, cs;
    sincospif(-2.0f * (float)(((uint32_t)n2 * (uint32_t)t) & nmask) / (float)(1u << log2N), &sn, &cs);
    const float2 wbase = make_float2(cs, sn);
    sincospif(-2.0f * (float)(((uint32_t)n2 * (uint32_t)T) & nmask) / (float)(1u << log2N), &sn, &cs);
    const float2 wstep = make_float2(cs, sn);
    float2 p[E];
    p[0] = wbase;
    p[1] = cmul(wbase, wstep);
    // p[e] = wbase * wstep^e with a log-depth tree on the powers of wstep
    {
        float2 q[E];
        q[0] = make_float2(1.0f, 0.0f);
        q[1] = wstep;
#pragma unroll
        for (int e = 2; e < E; e++) q[e] = cmul(q[e / 2], q[e - e / 2]);
#pragma unroll
        for (int e = 2; e < E; e++) p[e] = cmul(wbase, q[e]);
    }
    float2* __restrict__ out = a.inter + (size_t)f * ((size_t)1 << log2N);
#pragma unroll
    for (int e = 0; e < E; e++) {
        const int k1 = t + T * e;
        out[(size_t)k1 * N2 + n2] = cmul(v[e], p[e]);
    }
}

// ---------------------------------------------------------------------------------------------
// rows kernel: grid (N1/B, frames), block T*B, thread (b,t) with t fastest (contiguous rows).
// FROM_SAMPLES: N1 == 1, the rows are the windowed frames themselves (N <= 4096).
// ---------------------------------------------------------------------------------------------
template <class P, int B, bool FROM_SAMPLES>
__global__ void __launch_bounds__(P::T* B)
fft_rows_kernel(SpectrumArgs a, int N1, int log2N) {
    extern __shared__ __align__(16) float2 sm[];
    constexpr int E = P::E, T = P::T, L = P::L;
    const int t = threadIdx.x % T, b = threadIdx.x / T;
    float2 v[E];

    if constexpr (FROM_SAMPLES) {
        const int f = blockIdx.x * B + b;
        const bool live = f < a.frames;
        const uint32_t base = a.start + (uint32_t)f * a.frame_stride;
        const float2* __restrict__ in = reinterpret_cast<const float2*>(a.in);
#pragma unroll
        for (int e = 0; e < E; e++) {
            const int n = t + T * e;
            if (live && n < a.nz) {
                const float2 x = __ldg(in + ((base + (uint32_t)n) & a.ring_mask));
                const float w = __ldg(a.window + n);
                v[e] = make_float2(x.x * w, x.y * w);
            } else {
                v[e] = make_float2(0.0f, 0.0f);
            }
        }
        block_fft<P, false, B>(v, sm, t, b);
        if (live) {
#pragma unroll
            for (int e = 0; e < E; e++) {
                const int k = t + T * e;
                if (a.rows) a.rows[(size_t)f * L + k] = power_db(v[e]);
                if (a.X) a.X[(size_t)f * L + k] = v[e];
            }
        }
    } else {
        const int f = blockIdx.y;
        const int k1_0 = blockIdx.x * B;
        const size_t N = (size_t)1 << log2N;
        const float2* __restrict__ A = a.inter + (size_t)f * N + (size_t)(k1_0 + b) * L;
#pragma unroll
        for (int e = 0; e < E; e++) v[e] = A[t + T * e];
        block_fft<P, false, B>(v, sm, t, b);
        if (a.X) {
#pragma unroll
            for (int e = 0; e < E; e++) a.X[(size_t)f * N + (size_t)(k1_0 + b) + (size_t)N1 * (t + T * e)] = v[e];
        }
        if (a.rows) {
            // transpose through shared memory so each k2 writes B contiguous floats
            float* so = reinterpret_cast<float*>(sm);
#pragma unroll
            for (int e = 0; e < E; e++) so[(t + T * e) * (B + 1) + b] = power_db(v[e]);
            __syncthreads();
            float* __restrict__ row = a.rows + (size_t)f * N + k1_0;
            for (int i = threadIdx.x; i < L * B; i += T * B) {
                const int k2 = i / B, bb = i % B;
                row[(size_t)N1 * k2 + bb] = so[k2 * (B + 1) + bb];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Host dispatch
// ---------------------------------------------------------------------------------------------
using P64 = FftPlan<64, 8, 8, 8, 1>;
using P128 = FftPlan<128, 16, 16, 8, 1>;
using P256 = FftPlan<256, 16, 16, 16, 1>;
using P512 = FftPlan<512, 32, 32, 16, 1>;
using P1024 = FftPlan<1024, 32, 32, 32, 1>;
using P2048 = FftPlan<2048, 16, 16, 16, 8>;
using P4096 = FftPlan<4096, 16, 16, 16, 16>;

template <class P, int B>
static cudaError_t launch_cols(const SpectrumArgs& a, int N2, int log2N, cudaStream_t st) {
    constexpr size_t smem = fft_smem_bytes<P, true, B>();
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(fft_cols_kernel<P, B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    dim3 grid(N2 / B, a.frames);
    fft_cols_kernel<P, B><<<grid, P::T * B, smem, st>>>(a, N2, log2N);
    return cudaGetLastError();
}

template <class P, int B, bool FROM_SAMPLES>
static cudaError_t launch_rows(const SpectrumArgs& a, int N1, int log2N, cudaStream_t st) {
    constexpr size_t ex = fft_smem_bytes<P, false, B>();
    constexpr size_t tr = FROM_SAMPLES ? 0 : (size_t)P::L * (B + 1) * sizeof(float);
    constexpr size_t smem = ex > tr ? ex : tr;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(fft_rows_kernel<P, B, FROM_SAMPLES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_done = true;
    }
    dim3 grid(FROM_SAMPLES ? ceil_div(a.frames, B) : N1 / B, FROM_SAMPLES ? 1 : a.frames);
    fft_rows_kernel<P, B, FROM_SAMPLES><<<grid, P::T * B, smem, st>>>(a, N1, log2N);
    return cudaGetLastError();
}

int spectrum_split(int N, int* N1, int* N2) {
    int lg = 0;
    while ((1 << lg) < N) lg++;
    if ((1 << lg) != N || lg < 6 || lg > 22) return -1;
    if (lg <= 12) { *N1 = 1; *N2 = N; return lg; }
    const int l1 = lg / 2;          // N1 <= N2
    *N1 = 1 << l1; *N2 = 1 << (lg - l1);
    return lg;
}

cudaError_t launch_spectrum(int N, const SpectrumArgs& a, cudaStream_t st, long long* launches) {
    int N1, N2;
    const int lg = spectrum_split(N, &N1, &N2);
    if (lg < 0 || a.frames <= 0) return cudaErrorInvalidValue;
    cudaError_t e = cudaSuccess;
    if (N1 == 1) {
        switch (N) {
        case 64: e = launch_rows<P64, 16, true>(a, 1, lg, st); break;
        case 128: e = launch_rows<P128, 16, true>(a, 1, lg, st); break;
        case 256: e = launch_rows<P256, 8, true>(a, 1, lg, st); break;
        case 512: e = launch_rows<P512, 8, true>(a, 1, lg, st); break;
        case 1024: e = launch_rows<P1024, 4, true>(a, 1, lg, st); break;
        case 2048: e = launch_rows<P2048, 2, true>(a, 1, lg, st); break;
        case 4096: e = launch_rows<P4096, 1, true>(a, 1, lg, st); break;
        }
        if (launches) *launches += 1;
        return e;
    }
    switch (N1) {
    case 64: e = launch_cols<P64, 16>(a, N2, lg, st); break;
    case 128: e = launch_cols<P128, 16>(a, N2, lg, st); break;
    case 256: e = launch_cols<P256, 16>(a, N2, lg, st); break;
    case 512: e = launch_cols<P512, 16>(a, N2, lg, st); break;
    case 1024: e = launch_cols<P1024, 8>(a, N2, lg, st); break;
    case 2048: e = launch_cols<P2048, 8>(a, N2, lg, st); break;
    default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    switch (N2) {
    case 128: e = launch_rows<P128, 16, false>(a, N1, lg, st); break;
    case 256: e = launch_rows<P256, 16, false>(a, N1, lg, st); break;
    case 512: e = launch_rows<P512, 16, false>(a, N1, lg, st); break;
    case 1024: e = launch_rows<P1024, 16, false>(a, N1, lg, st); break;
    case 2048: e = launch_rows<P2048, 8, false>(a, N1, lg, st); break;
    default: return cudaErrorInvalidValue;
    }
    if (launches) *launches += 2;
    return e;
}

} // namespace sdrpp
