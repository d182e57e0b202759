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
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>
#include "../../include/sdrpp_cuda.h"

#ifndef SDRPP_FFT_CB
#define SDRPP_FFT_CB 8
#endif
#ifndef SDRPP_FFT_RB
#define SDRPP_FFT_RB 8
#endif

namespace sdrpp {

__host__ __device__ constexpr int ilog2c(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

__device__ __forceinline__ float power_db(float2 x) {
    // volk_32fc_s32f_power_spectrum_32f with norm 1: log2(re^2+im^2) * 10/log2(10), -inf -> -127*3.0103
    float l = __log2f(x.x * x.x + x.y * x.y);
    if (isinf(l)) l = copysignf(127.0f, l);
    return 3.01029995663981209120f * l;
}

// ---------------------------------------------------------------------------------------------
// Twiddle tables: exp(-2*pi*i*j/L) computed in double on the host, rounded once. One set per (device, N).
// ---------------------------------------------------------------------------------------------
struct SpectrumTables {
    float2* tw1 = nullptr;  // L = N1 (cols transform; also the coarse factor of the four-step twiddle)
    float2* tw2 = nullptr;  // L = N2 (rows transform)
    float2* twlo = nullptr; // exp(-2*pi*i*j/N), j < N2 (fine factor of the four-step twiddle)
};
static std::mutex g_tab_mtx;
static std::map<std::pair<int, int>, SpectrumTables> g_tabs;

static cudaError_t upload_table(float2** dst, int count, double denom) {
    std::vector<float2> h((size_t)count);
    for (int j = 0; j < count; j++) {
        const double a = -2.0 * 3.14159265358979323846 * (double)j / denom;
        h[(size_t)j] = make_float2((float)cos(a), (float)sin(a));
    }
    cudaError_t e = cudaMalloc((void**)dst, sizeof(float2) * (size_t)count);
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(*dst, h.data(), sizeof(float2) * (size_t)count, cudaMemcpyHostToDevice);
    return e == cudaSuccess ? cudaStreamSynchronize(cudaStreamLegacy) : e; // complete before any non-blocking stream reads it
}

static cudaError_t get_tables(int N, int N1, int N2, SpectrumTables* out) {
    std::lock_guard<std::mutex> lck(g_tab_mtx);
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    auto key = std::make_pair(dev, N);
    auto it = g_tabs.find(key);
    if (it == g_tabs.end()) {
        SpectrumTables t;
        if ((e = upload_table(&t.tw2, N2, (double)N2)) != cudaSuccess) return e;
        if (N1 > 1) {
            if ((e = upload_table(&t.tw1, N1, (double)N1)) != cudaSuccess) return e;
            if ((e = upload_table(&t.twlo, N2, (double)N)) != cudaSuccess) return e;
        }
        it = g_tabs.emplace(key, t).first;
    }
    *out = it->second;
    return cudaSuccess;
}

// Twiddle tables travel to shared memory by cp.async (16 B per copy, no register round trip): the copies are in flight
// beside the frame's own loads instead of in front of the transform. As a load -> store loop every trip was one more
// serialised L2 / DRAM round trip ahead of the first butterfly (ncu r2t: long_scoreboard 65 % of the cols kernel's samples).
// n is a multiple of 2 (every table length is a power of two >= 64) and both sides are 16-byte aligned.
__device__ __forceinline__ void stage_table(float2* dst, const float2* __restrict__ src, int n) {
    for (int i = threadIdx.x * 2; i < n; i += blockDim.x * 2)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + i)), "l"(src + i) : "memory");
}
// (the matching cp.async.wait_all sits in front of block_fft's first __syncthreads, fft_core.cuh)

// ---------------------------------------------------------------------------------------------
// cols kernel: grid (N2/B, frames), block T*B, thread (t,b) with b fastest (coalesced columns)
// ---------------------------------------------------------------------------------------------
// SQUARE: N1 == N2 == L, the frame is fully windowed (nz == N) and does not wrap in the ring -- the saturated 1M-point
// waterfall. Every stride is then a compile-time constant and the 32 frame loads, window loads and stores of a thread are one
// base register plus an immediate each (a third of the general kernel's instructions were address arithmetic and range
// predicates, SASS r2t).
template <class P, int B, bool SQUARE>
__device__ __forceinline__ void fft_cols_body(const SpectrumArgs& a, const SpectrumTables& tabs, float2* sm, int N2r, int log2N, int log2N2r) {
    constexpr int E = P::E, T = P::T, L = P::L;
    const int N2 = SQUARE ? L : N2r;
    float2* tw = sm;              // [L]   exp(-2 pi i j / N1)
    float2* twlo = sm + L;        // [N2]  exp(-2 pi i j / N)
    float2* ex = twlo + N2;       // exchange buffer
    const int b = threadIdx.x % B, t = threadIdx.x / B;
    const int n2 = blockIdx.x * B + b;
    const int f = blockIdx.y;
    const uint32_t base = a.start + (uint32_t)f * a.frame_stride;
    const float2* __restrict__ in = reinterpret_cast<const float2*>(a.in);

    stage_table(tw, tabs.tw1, L);
    stage_table(twlo, tabs.twlo, N2);
    float2 v[E];
    if constexpr (SQUARE) {
        const uint32_t first = base & a.ring_mask;
        const float* __restrict__ w = a.window + (t * L + n2);
#ifndef SDRPP_FFT_NO_STAGE_X
        // The CTA's tile of the frame (L rows of B samples) goes to shared memory by 16-byte cp.async (no registers held, no
        // L1 lines: the copies in flight are not bounded by the 60 KB of L1), into the exchange buffer, which is free until the
        // first butterflies are done. Needs an even frame start (two samples per copy). 18 frames of 1M points 113.3 -> 123.9 GS/s
        // on one box, a single frame unchanged (9.86 us, ncu cold; profiles/r2x_fft_ab.txt).
        const bool staged = (first & 1u) == 0 && (reinterpret_cast<uintptr_t>(in) & 15) == 0;
        if (staged) {
            constexpr int CPR = B / 2;                               // 16-byte copies per row of the tile
            const float2* __restrict__ xf = in + first + blockIdx.x * B;
            float2* xs = ex;
#pragma unroll
            for (int i = 0; i < (L * CPR) / (T * B); i++) {
                const int c = threadIdx.x + i * (T * B);
                const int row = c / CPR, part = c % CPR;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(xs + row * B + part * 2)),
                             "l"(xf + (size_t)row * L + part * 2) : "memory");
            }
            float wv[E];
#pragma unroll
            for (int e = 0; e < E; e++) wv[e] = __ldg(w + e * (T * L));
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
#pragma unroll
            for (int e = 0; e < E; e++) {
                const float2 xv = xs[(t + T * e) * B + b];
                v[e] = make_float2(xv.x * wv[e], xv.y * wv[e]);
            }
            __syncthreads();                                         // the first exchange overwrites the tile
        } else
#endif
        {
            const float2* __restrict__ x = in + first + (t * L + n2);
#pragma unroll
            for (int e = 0; e < E; e++) {
                const float2 xv = __ldg(x + e * (T * L));
                const float wv = __ldg(w + e * (T * L));
                v[e] = make_float2(xv.x * wv, xv.y * wv);
            }
        }
    } else {
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
    }
    block_fft<P, true, B>(v, ex, tw, t, b); // the first exchange's __syncthreads also publishes the tables

    // four-step twiddle W_N^m, m = n2*k1 < N: W_N^m = exp(-2 pi i (m >> log2N2) / N1) * exp(-2 pi i (m & (N2-1)) / N)
    const int log2N2 = SQUARE ? ilog2c(L) : log2N2r;
    float2* __restrict__ out = a.inter + (size_t)f * ((size_t)1 << log2N) + (size_t)t * N2 + n2;
    const uint32_t lomask = (uint32_t)N2 - 1u;
    uint32_t m = (uint32_t)n2 * (uint32_t)t;
    const uint32_t dm = (uint32_t)n2 * (uint32_t)T;
#pragma unroll
    for (int e = 0; e < E; e++) {
        const float2 w = cmul(tw[m >> log2N2], twlo[m & lomask]);
        out[(size_t)e * T * N2] = cmul(v[e], w);       // row k1 = t + T*e
        m += dm;
    }
}

template <class P, int B>
__global__ void __launch_bounds__(P::T* B)
fft_cols_kernel(const SpectrumArgs* __restrict__ ap, SpectrumTables tabs, int N2, int log2N, int log2N2) {
    const SpectrumArgs a = *ap;   // per-block arguments (launcher descriptor)
    if ((int)blockIdx.y >= a.frames) return;
    extern __shared__ __align__(16) float2 sm[];
    const uint32_t first = (a.start + blockIdx.y * a.frame_stride) & a.ring_mask;
    const bool square = N2 == P::L && a.nz == (1 << log2N) && (unsigned long long)first + (1ull << log2N) <= (unsigned long long)a.ring_mask + 1ull;
    if (square) fft_cols_body<P, B, true>(a, tabs, sm, N2, log2N, log2N2);
    else fft_cols_body<P, B, false>(a, tabs, sm, N2, log2N, log2N2);
}

// ---------------------------------------------------------------------------------------------
// rows kernel: grid (N1/B, frames), block T*B, thread (b,t) with t fastest (contiguous rows).
// FROM_SAMPLES: N1 == 1, the rows are the windowed frames themselves (N <= 4096).
// ---------------------------------------------------------------------------------------------
template <class P, int B, bool FROM_SAMPLES>
__global__ void __launch_bounds__(P::T* B)
fft_rows_kernel(const SpectrumArgs* __restrict__ ap, SpectrumTables tabs, int N1, int log2N) {
    const SpectrumArgs a = *ap;   // per-block arguments (launcher descriptor)
    if (!FROM_SAMPLES && (int)blockIdx.y >= a.frames) return;
    extern __shared__ __align__(16) float2 sm[];
    constexpr int E = P::E, T = P::T, L = P::L;
    float2* tw = sm;        // [L] exp(-2 pi i j / N2)
    float2* ex = sm + L;    // exchange buffer, later the transposed dB tile
    const int t = threadIdx.x % T, b = threadIdx.x / T;
    float2 v[E];

    if constexpr (FROM_SAMPLES) {
        const int f = blockIdx.x * B + b;
        const bool live = f < a.frames;
        const uint32_t base = a.start + (uint32_t)f * a.frame_stride;
        const float2* __restrict__ in = reinterpret_cast<const float2*>(a.in);
        stage_table(tw, tabs.tw2, L);
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
            block_fft<P, false, B>(v, ex, tw, t, b);
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
        stage_table(tw, tabs.tw2, L);
#pragma unroll
        for (int e = 0; e < E; e++) v[e] = A[t + T * e];
            block_fft<P, false, B>(v, ex, tw, t, b);
        if (a.X) {
#pragma unroll
            for (int e = 0; e < E; e++) a.X[(size_t)f * N + (size_t)(k1_0 + b) + (size_t)N1 * (t + T * e)] = v[e];
        }
        if (a.rows) {
            // transpose through shared memory so each k2 writes B contiguous floats
            float* so = reinterpret_cast<float*>(ex);
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
#ifdef SDRPP_FFT1024_E16
using P1024 = FftPlan<1024, 16, 16, 16, 4>;
#else
using P1024 = FftPlan<1024, 32, 32, 32, 1>;
#endif
using P2048 = FftPlan<2048, 16, 16, 16, 8>;
using P4096 = FftPlan<4096, 16, 16, 16, 16>;

// Shared-memory carve-out preference of the spectrum kernels: the driver's own choice from the kernel's occupancy (two 84 KB
// CTAs per SM -> the 196 KB setting, 60 KB of L1), not the largest setting every other kernel of the library asks for
// (common.cuh): these kernels read their frame with per-thread loads of 64-B row segments, and the misses they can keep in
// flight scale with the L1 that is left. One box, max -> default (profiles/r2x_fft_ab.txt): 18 frames of 1M points
// 89.6 -> 113.2 GS/s, one frame 12.4 + 8.7 -> 9.8 + 8.4 us (ncu, cold), the streaming bench line unchanged (6831 / 6806
// MS/s: the spectrum takes turns with the persistent stage-1 CTAs either way). SDRPP_FFT_CARVEOUT = max | default | percent.
static int fft_carveout() {
    static const int v = [] {
        const char* e = getenv("SDRPP_FFT_CARVEOUT");
        if (!e || !*e || !strcmp(e, "default")) return (int)cudaSharedmemCarveoutDefault;
        if (!strcmp(e, "max")) return (int)cudaSharedmemCarveoutMaxShared;
        return atoi(e);
    }();
    return v;
}

static int ilog2(int v) { int l = 0; while ((1 << l) < v) l++; return l; }

template <class P, int B>
static cudaError_t launch_cols(Launcher& L, int sid, const SpectrumArgs& a, const SpectrumArgs* d_a, const SpectrumTables& tabs, int N2, int log2N) {
    const size_t smem = (P::L + (size_t)N2 + fft_exchange_elems<P, true, B>()) * sizeof(float2);
    if (cudaError_t e = ensure_dynamic_smem((const void*)fft_cols_kernel<P, B>, smem, fft_carveout()); e != cudaSuccess) return e;
    dim3 grid(N2 / B, a.frames);
    return L.kernel(sid, (const void*)fft_cols_kernel<P, B>, grid, dim3(P::T * B), smem, d_a, tabs, N2, log2N, ilog2(N2));
}

template <class P, int B, bool FROM_SAMPLES>
static cudaError_t launch_rows(Launcher& L, int sid, const SpectrumArgs& a, const SpectrumArgs* d_a, const SpectrumTables& tabs, int N1, int log2N) {
    constexpr size_t ex = fft_exchange_elems<P, false, B>() * sizeof(float2);
    constexpr size_t tr = FROM_SAMPLES ? 0 : (size_t)P::L * (B + 1) * sizeof(float);
    constexpr size_t smem = P::L * sizeof(float2) + (ex > tr ? ex : tr);
    if (cudaError_t e = ensure_dynamic_smem((const void*)fft_rows_kernel<P, B, FROM_SAMPLES>, smem, fft_carveout()); e != cudaSuccess) return e;
    dim3 grid(FROM_SAMPLES ? ceil_div(a.frames, B) : N1 / B, FROM_SAMPLES ? 1 : a.frames);
    return L.kernel(sid, (const void*)fft_rows_kernel<P, B, FROM_SAMPLES>, grid, dim3(P::T * B), smem, d_a, tabs, N1, log2N);
}

// ---------------------------------------------------------------------------------------------
// Waterfall zoom: one warp per output pixel, lanes stride over the pixel's bin range, shuffle max.
// std::max(a, b) = (a < b) ? b : a, so a NaN bin never replaces the running maximum (as in doZoom).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fft_zoom_kernel(const float* __restrict__ rows, int N, const int* __restrict__ idx, int outSize, bool ranged, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int px = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (px >= outSize) return;
    const float* __restrict__ row = rows + (size_t)blockIdx.y * N;
    int i0 = idx[px];
    i0 = max(0, min(N - 1, i0));
    float m = row[i0];
    if (ranged) {
        const int i1 = min(N, idx[px + 1]);
        for (int j = i0 + 1 + lane; j < i1; j += 32) { const float v = row[j]; m = (m < v) ? v : m; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const float v = __shfl_xor_sync(0xffffffffu, m, o); m = (m < v) ? v : m; }
    }
    if (lane == 0) out[(size_t)blockIdx.y * outSize + px] = m;
}

cudaError_t launch_fft_zoom(Launcher& L, int sid, const float* rows, int N, int nrows, const int* idx, int outSize, bool ranged, float* out) {
    if (nrows <= 0 || outSize <= 0) return cudaSuccess;
    dim3 grid(ceil_div(outSize, 8), nrows);
    return L.kernel(sid, (const void*)fft_zoom_kernel, grid, dim3(256), 0, rows, N, idx, outSize, ranged, out);
}

// ---------------------------------------------------------------------------------------------
// WaterFall::calculateVFOSignalInfo (gui/widgets/waterfall.cpp:563-603) on the device, for every (VFO, row) pair of a
// block: strength = the largest bin inside the VFO's bandwidth, snr = strength - mean of the bins in the two half-
// bandwidth shoulders beside it (summed in double like the reference; the order of the additions differs). One CTA per
// pair; bins = (minSide, min, max, maxSide) are computed on the host with the reference's clamp arithmetic.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
signal_info_kernel(const float* __restrict__ rows, int N, const int4* __restrict__ bins, float2* __restrict__ out, int nsig) {
    __shared__ double ssum[8];
    __shared__ float smax[8];
    __shared__ int scnt[8];
    const int v = blockIdx.x, r = blockIdx.y, tid = threadIdx.x;
    const int4 b = bins[v];
    const float* __restrict__ row = rows + (size_t)r * N;
    double sum = 0.0;
    int cnt = 0;
    for (int i = b.x + tid; i < b.y; i += 256) { sum += (double)row[i]; cnt++; }          // left shoulder  [minSide, min)
    for (int i = b.z + 1 + tid; i < b.w; i += 256) { sum += (double)row[i]; cnt++; }      // right shoulder (max, maxSide)
    float m = -INFINITY;
    for (int i = b.y + tid; i <= b.z && i < N; i += 256) { const float x = row[i]; m = (x > m) ? x : m; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        const float x = __shfl_xor_sync(0xffffffffu, m, o); m = (x > m) ? x : m;
    }
    if ((tid & 31) == 0) { ssum[tid >> 5] = sum; smax[tid >> 5] = m; scnt[tid >> 5] = cnt; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; w++) { sum += ssum[w]; cnt += scnt[w]; m = (smax[w] > m) ? smax[w] : m; }
        const double avg = sum / (double)cnt;                    // 0/0 = NaN when there is no shoulder, as in the reference
        out[(size_t)r * nsig + v] = make_float2(m, (float)((double)m - avg));   // strength = max; snr = max - avg (float - double)
    }
}

cudaError_t launch_signal_info(Launcher& L, int sid, const float* rows, int N, int nrows, const int4* bins, int nsig, float2* out) {
    if (nrows <= 0 || nsig <= 0) return cudaSuccess;
    return L.kernel(sid, (const void*)signal_info_kernel, dim3((unsigned)nsig, (unsigned)nrows), dim3(256), 0, rows, N, bins, out, nsig);
}

// ---------------------------------------------------------------------------------------------
// The waterfall's per-line display state on the zoomed row (WaterFall::pushFFT, gui/widgets/waterfall.cpp:918-956):
// FFT smoothing latest = alpha*latest + beta*smoothingBuf (three VOLK calls: two multiplies and an add, each rounded)
// and peak hold hold[i] = max(latest[i], hold[i] - speed) for i >= 1. Sequential from line to line, parallel over the
// pixels: one thread per pixel walks the block's rows in order. `latest` keeps the last (smoothed) row, which is what
// setFFTSmoothing(true) seeds the smoothing buffer with (waterfall.cpp:1197-1201).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fft_display_kernel(float* __restrict__ zoom, int W, int nrows, bool smoothing, float alpha, float beta, float* __restrict__ smooth,
                   bool hold_on, float hold_speed, float* __restrict__ hold, float* __restrict__ latest) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= W) return;
    float sb = smooth[i], hb = hold[i], v = latest[i];
    for (int r = 0; r < nrows; r++) {
        v = zoom[(size_t)r * W + i];
        if (smoothing) {
            sb = __fadd_rn(__fmul_rn(sb, beta), __fmul_rn(v, alpha));
            v = sb;
            zoom[(size_t)r * W + i] = v;
        }
        if (hold_on && i >= 1) { const float d = __fsub_rn(hb, hold_speed); hb = (v < d) ? d : v; }   // std::max(latest, hold - speed)
    }
    smooth[i] = sb; hold[i] = hb; latest[i] = v;
}

cudaError_t launch_fft_display(Launcher& L, int sid, float* zoom, int W, int nrows, bool smoothing, float alpha, float* smooth, bool hold_on,
                               float hold_speed, float* hold, float* latest) {
    if (nrows <= 0 || W <= 0) return cudaSuccess;
    const float beta = 1.0f - alpha;
    return L.kernel(sid, (const void*)fft_display_kernel, dim3((unsigned)ceil_div(W, 256)), dim3(256), 0, zoom, W, nrows, smoothing, alpha, beta, smooth,
                    hold_on, hold_speed, hold, latest);
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

static cudaError_t launch_spectrum_frames(Launcher& L, int sid, int N, int N1, int N2, int lg, const SpectrumTables& tabs, const SpectrumArgs& a, long long* launches) {
    cudaError_t e = cudaSuccess;
    const SpectrumArgs* d_a = L.push(a);
    if (!d_a) return cudaErrorMemoryAllocation;
    if (N1 == 1) {
        switch (N) {
        case 64: e = launch_rows<P64, 16, true>(L, sid, a, d_a, tabs, 1, lg); break;
        case 128: e = launch_rows<P128, 16, true>(L, sid, a, d_a, tabs, 1, lg); break;
        case 256: e = launch_rows<P256, 8, true>(L, sid, a, d_a, tabs, 1, lg); break;
        case 512: e = launch_rows<P512, 8, true>(L, sid, a, d_a, tabs, 1, lg); break;
        case 1024: e = launch_rows<P1024, 4, true>(L, sid, a, d_a, tabs, 1, lg); break;
        case 2048: e = launch_rows<P2048, 2, true>(L, sid, a, d_a, tabs, 1, lg); break;
        case 4096: e = launch_rows<P4096, 1, true>(L, sid, a, d_a, tabs, 1, lg); break;
        }
        if (launches) *launches += 1;
        return e;
    }
    switch (N1) {
    case 64: e = launch_cols<P64, 16>(L, sid, a, d_a, tabs, N2, lg); break;
    case 128: e = launch_cols<P128, 16>(L, sid, a, d_a, tabs, N2, lg); break;
    case 256: e = launch_cols<P256, 16>(L, sid, a, d_a, tabs, N2, lg); break;
    case 512: e = launch_cols<P512, 16>(L, sid, a, d_a, tabs, N2, lg); break;
    case 1024: e = launch_cols<P1024, SDRPP_FFT_CB>(L, sid, a, d_a, tabs, N2, lg); break;
    case 2048: e = launch_cols<P2048, 8>(L, sid, a, d_a, tabs, N2, lg); break;
    default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    switch (N2) {
    case 128: e = launch_rows<P128, 16, false>(L, sid, a, d_a, tabs, N1, lg); break;
    case 256: e = launch_rows<P256, 16, false>(L, sid, a, d_a, tabs, N1, lg); break;
    case 512: e = launch_rows<P512, 16, false>(L, sid, a, d_a, tabs, N1, lg); break;
    case 1024: e = launch_rows<P1024, SDRPP_FFT_RB, false>(L, sid, a, d_a, tabs, N1, lg); break;
    case 2048: e = launch_rows<P2048, 8, false>(L, sid, a, d_a, tabs, N1, lg); break;
    default: return cudaErrorInvalidValue;
    }
    if (launches) *launches += 2;
    return e;
}

// All frames of a call go through one pair of launches. Splitting a large batch into chunks whose four-step intermediate stays
// in the L2 (32 MB at a time, the same region reused) was measured and is slower: 18 / 64 frames of 1M points 86.6 / 91.0 ->
// 74.6 / 75.7 GS/s, 256 frames of 64K 131 -> 113 GS/s (profiles/r2x_fft_ab.txt) -- the partial last wave of every small launch
// costs more than the intermediate's trip through HBM.
cudaError_t launch_spectrum(Launcher& L, int sid, int N, const SpectrumArgs& a, long long* launches) {
    int N1, N2;
    const int lg = spectrum_split(N, &N1, &N2);
    if (lg < 0 || a.frames <= 0) return cudaErrorInvalidValue;
    SpectrumTables tabs;
    cudaError_t e = get_tables(N, N1, N2, &tabs);
    if (e != cudaSuccess) return e;
    return launch_spectrum_frames(L, sid, N, N1, N2, lg, tabs, a, launches);
}

} // namespace sdrpp
