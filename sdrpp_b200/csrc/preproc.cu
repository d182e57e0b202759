// Ingest and pre-processing kernels: source sample conversion, the front-end
// PowerDecimator stages (A3/A4), the DC blocker (A5) and conjugation (A6). All of them write
// cf32 into a ring (index & mask); the last enabled one writes the device IQ ring that the
// spectrum and channelizer kernels read.
#include "common.cuh"
#include "kernels.h"

namespace sdrpp {

__device__ __forceinline__ void ring_store(RingRef dst, uint32_t idx, float2 v, bool conj) {
    if (conj) v.y = -v.y;
    dst.base[idx & dst.mask] = v;
}

// ---------------------------------------------------------------------------------------------
// Conversion. HBM-bound: 2/4/8 B in, 8 B out per sample. Each thread converts 4 consecutive
// samples from one 8/16/32-byte load and, when the destination is aligned, two 16-byte stores.
// ---------------------------------------------------------------------------------------------
struct IngestArgs {      // per-block arguments, read through the launcher's descriptor
    const void* raw;
    int count;
    RingRef dst;
    uint32_t pos;
    int conj, vec_ok;
    float scale;
};

template <int FMT>
__global__ void __launch_bounds__(256)
ingest_kernel(const IngestArgs* __restrict__ ap) {
    const void* __restrict__ raw = ap->raw;
    const int count = ap->count;
    const RingRef dst = ap->dst;
    const uint32_t pos = ap->pos;
    const bool conj = ap->conj != 0, vec_ok = ap->vec_ok != 0;
    const float scale = ap->scale;
    const int stride = gridDim.x * blockDim.x;
    const int nquads = count >> 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += stride) {
        float2 s[4];
        if (vec_ok && (FMT <= 5 || FMT >= 9)) {
            if constexpr (FMT == 0) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(raw) + 2 * q);
                const float4 b = __ldg(reinterpret_cast<const float4*>(raw) + 2 * q + 1);
                s[0] = make_float2(a.x, a.y); s[1] = make_float2(a.z, a.w);
                s[2] = make_float2(b.x, b.y); s[3] = make_float2(b.z, b.w);
            } else if constexpr (FMT == 1 || FMT == 2) {
                const uint2 w = __ldg(reinterpret_cast<const uint2*>(raw) + q);
                const unsigned int b[8] = { w.x & 255u, (w.x >> 8) & 255u, (w.x >> 16) & 255u, w.x >> 24,
                                            w.y & 255u, (w.y >> 8) & 255u, (w.y >> 16) & 255u, w.y >> 24 };
#pragma unroll
                for (int i = 0; i < 4; i++)
                    s[i] = FMT == 1 ? make_float2(cvt_u8_rtl(b[2 * i]), cvt_u8_rtl(b[2 * i + 1]))
                                    : make_float2(cvt_u8_tcp(b[2 * i]), cvt_u8_tcp(b[2 * i + 1]));
            } else if constexpr (FMT == 3 || FMT == 9) {
                const uint2 w = __ldg(reinterpret_cast<const uint2*>(raw) + q);
                const int b[8] = { (int)(signed char)(w.x & 255u), (int)(signed char)((w.x >> 8) & 255u),
                                   (int)(signed char)((w.x >> 16) & 255u), (int)(signed char)(w.x >> 24),
                                   (int)(signed char)(w.y & 255u), (int)(signed char)((w.y >> 8) & 255u),
                                   (int)(signed char)((w.y >> 16) & 255u), (int)(signed char)(w.y >> 24) };
#pragma unroll
                for (int i = 0; i < 4; i++)
                    s[i] = FMT == 3 ? make_float2(cvt_i8(b[2 * i]), cvt_i8(b[2 * i + 1]))
                                    : make_float2(__fdiv_rn((float)b[2 * i], scale), __fdiv_rn((float)b[2 * i + 1], scale));
            } else if constexpr (FMT == 4 || FMT == 5 || FMT == 10) {
                const uint4 w = __ldg(reinterpret_cast<const uint4*>(raw) + q);
                const unsigned int u[4] = { w.x, w.y, w.z, w.w };
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int re = (int)(short)(u[i] & 0xFFFFu), im = (int)(short)(u[i] >> 16);
                    s[i] = FMT == 4 ? make_float2(cvt_i16_file(re), cvt_i16_file(im))
                         : FMT == 5 ? make_float2(cvt_i16_volk(re), cvt_i16_volk(im))
                                    : make_float2(__fdiv_rn((float)re, scale), __fdiv_rn((float)im, scale));
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) s[i] = load_sample<FMT>(raw, (size_t)4 * q + i, scale);
        }
        if (conj) {
#pragma unroll
            for (int i = 0; i < 4; i++) s[i].y = -s[i].y;
        }
        const uint32_t i0 = (pos + 4u * (uint32_t)q) & dst.mask;
        if ((i0 & 1u) == 0 && i0 <= dst.mask - 3u) {
            float4* o = reinterpret_cast<float4*>(dst.base + i0);
            o[0] = make_float4(s[0].x, s[0].y, s[1].x, s[1].y);
            o[1] = make_float4(s[2].x, s[2].y, s[3].x, s[3].y);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) dst.base[(pos + 4u * (uint32_t)q + i) & dst.mask] = s[i];
        }
    }
    // tail samples
    const int tail0 = nquads << 2;
    const int t = tail0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < count) {
        float2 v = load_sample<FMT>(raw, (size_t)t, scale);
        ring_store(dst, pos + (uint32_t)t, v, conj);
    }
}

cudaError_t launch_ingest(Launcher& L, int sid, int fmt, const void* raw, int count, RingRef dst, uint32_t pos, bool conj, float scale) {
    if (count <= 0) return cudaSuccess;
    const int threads = 256;
    int blocks = ceil_div(ceil_div(count, 4), threads);
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    IngestArgs a{ raw, count, dst, pos, conj ? 1 : 0, (((uintptr_t)raw & 15u) == 0 && !(fmt >= 6 && fmt <= 8)) ? 1 : 0, scale };
    const IngestArgs* d = L.push(a);
    if (!d) return cudaErrorMemoryAllocation;
    const void* fn = nullptr;
    switch (fmt) {
    case 0: fn = (const void*)ingest_kernel<0>; break;
    case 1: fn = (const void*)ingest_kernel<1>; break;
    case 2: fn = (const void*)ingest_kernel<2>; break;
    case 3: fn = (const void*)ingest_kernel<3>; break;
    case 4: fn = (const void*)ingest_kernel<4>; break;
    case 5: fn = (const void*)ingest_kernel<5>; break;
    case 6: fn = (const void*)ingest_kernel<6>; break;
    case 7: fn = (const void*)ingest_kernel<7>; break;
    case 8: fn = (const void*)ingest_kernel<8>; break;
    case 9: fn = (const void*)ingest_kernel<9>; break;
    case 10: fn = (const void*)ingest_kernel<10>; break;
    default: return cudaErrorInvalidValue;
    }
    return L.kernel(sid, fn, dim3((unsigned)blocks), dim3((unsigned)threads), 0, d);
}

// ---------------------------------------------------------------------------------------------
// SDR++ server wire compression (dsp::compression::SampleStreamCompressor::process,
// sample_stream_compressor.h:26-60): the packet's scaler is the largest SIGNED float of the block
// (volk_32f_index_max_32u, then in[maxIdx]), and every scalar becomes rintf(x * (128|32768)/max)
// saturated to the integer range (generic volk_32f_s32f_convert_8i / _16i). HBM-bound: 4 B in,
// 1-2 B out per scalar, two passes over the block (the second one is served by L2 below ~100 MB).
// ---------------------------------------------------------------------------------------------
// Order-preserving float -> uint key: key(a) < key(b)  <=>  a < b for non-NaN a, b.
__device__ __forceinline__ unsigned int float_key(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned int k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__global__ void __launch_bounds__(256)
pcm_max_kernel(const float* __restrict__ in, int nscalars, unsigned int* __restrict__ key_out) {
    unsigned int best = 0u;
    const int stride = gridDim.x * blockDim.x;
    const int nquads = nscalars >> 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(in) + q);
        best = max(best, max(max(float_key(v.x), float_key(v.y)), max(float_key(v.z), float_key(v.w))));
    }
    const int t = (nquads << 2) + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nscalars) best = max(best, float_key(__ldg(in + t)));
    best = __reduce_max_sync(0xFFFFFFFFu, best);
    __shared__ unsigned int s_best[8];
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        best = threadIdx.x < (blockDim.x >> 5) ? s_best[threadIdx.x] : 0u;
        best = __reduce_max_sync(0xFFFFFFFFu, best);
        if (threadIdx.x == 0) atomicMax(key_out, best);
    }
}

template <int BITS>
__device__ __forceinline__ int pcm_quantise(float x, float scalar) {
    const float r = __fmul_rn(x, scalar);
    constexpr float lo = BITS == 8 ? -128.0f : -32768.0f, hi = BITS == 8 ? 127.0f : 32767.0f;
    // the comparisons are false for NaN, like the reference's; the cast of a NaN is then 0 here
    return r > hi ? (int)hi : r < lo ? (int)lo : __float2int_rn(r);
}

template <int BITS>
__global__ void __launch_bounds__(256)
pcm_pack_kernel(const float* __restrict__ in, int nscalars, const unsigned int* __restrict__ key, void* __restrict__ out,
                float* __restrict__ scaler_out) {
    const float maxv = key_float(*key);
    const float scalar = __fdiv_rn(BITS == 8 ? 128.0f : 32768.0f, maxv);
    if (blockIdx.x == 0 && threadIdx.x == 0) *scaler_out = maxv;
    const int stride = gridDim.x * blockDim.x;
    const int nquads = nscalars >> 2;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(in) + q);
        const int a = pcm_quantise<BITS>(v.x, scalar), b = pcm_quantise<BITS>(v.y, scalar);
        const int c = pcm_quantise<BITS>(v.z, scalar), d = pcm_quantise<BITS>(v.w, scalar);
        if constexpr (BITS == 8)
            reinterpret_cast<unsigned int*>(out)[q] = (a & 255) | ((b & 255) << 8) | ((c & 255) << 16) | ((unsigned)(d & 255) << 24);
        else
            reinterpret_cast<uint2*>(out)[q] = make_uint2((a & 0xFFFF) | ((unsigned)(b & 0xFFFF) << 16), (c & 0xFFFF) | ((unsigned)(d & 0xFFFF) << 16));
    }
    const int t = (nquads << 2) + blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nscalars) {
        const int a = pcm_quantise<BITS>(__ldg(in + t), scalar);
        if constexpr (BITS == 8) reinterpret_cast<signed char*>(out)[t] = (signed char)a;
        else reinterpret_cast<short*>(out)[t] = (short)a;
    }
}

cudaError_t launch_pcm_compress(int bits, const float* in, int nscalars, unsigned int* key, void* out, float* scaler_out,
                                cudaStream_t st) {
    if (nscalars <= 0 || (bits != 8 && bits != 16)) return cudaErrorInvalidValue;
    int blocks = ceil_div(ceil_div(nscalars, 4), 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    cudaError_t e = cudaMemsetAsync(key, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return e;
    pcm_max_kernel<<<blocks, 256, 0, st>>>(in, nscalars, key);
    if (bits == 8) pcm_pack_kernel<8><<<blocks, 256, 0, st>>>(in, nscalars, key, out, scaler_out);
    else pcm_pack_kernel<16><<<blocks, 256, 0, st>>>(in, nscalars, key, out, scaler_out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Front-end decimating FIR stage (real taps). One output per thread; the input window of a warp
// is contiguous (32*D + T samples) and served by L1 after the first touch.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
decim_stage_kernel(const float2* __restrict__ buf, const float* __restrict__ taps, int T, int D, int offset,
                   int nout, RingRef dst, uint32_t pos, bool conj) {
    extern __shared__ float s_taps[];
    for (int k = threadIdx.x; k < T; k += blockDim.x) s_taps[k] = taps[k];
    __syncthreads();
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= nout) return;
    const float2* __restrict__ x = buf + offset + (size_t)m * D;
    float re = 0.0f, im = 0.0f;
    for (int k = 0; k < T; k++) {
        const float2 v = x[k];
        const float h = s_taps[k];
        re = fmaf(v.x, h, re);
        im = fmaf(v.y, h, im);
    }
    ring_store(dst, pos + (uint32_t)m, make_float2(re, im), conj);
}

cudaError_t launch_decim_stage(const float2* buf, const float* taps, int T, int D, int offset, int nout,
                               RingRef dst, uint32_t pos, bool conj, cudaStream_t st) {
    if (nout <= 0) return cudaSuccess;
    decim_stage_kernel<<<ceil_div(nout, 256), 256, T * sizeof(float), st>>>(buf, taps, T, D, offset, nout, dst, pos, conj);
    return cudaGetLastError();
}

__global__ void shift_history_kernel(float2* buf, int hist, int count) {
    // single CTA: read everything first, then write (source and destination may overlap)
    float2 v[4];
    int n = 0;
    for (int i = threadIdx.x; i < hist && n < 4; i += blockDim.x, n++) v[n] = buf[count + i];
    __syncthreads();
    n = 0;
    for (int i = threadIdx.x; i < hist && n < 4; i += blockDim.x, n++) buf[i] = v[n];
}

cudaError_t launch_shift_history(float2* buf, int hist, int count, cudaStream_t st) {
    if (hist <= 0 || count <= 0) return cudaSuccess;
    if (hist > 4096) return cudaErrorInvalidValue;
    shift_history_kernel<<<1, 1024, 0, st>>>(buf, hist, count);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// DC blocker. Reference recurrence (dc_blocker.h:54-60): out = in - off; off += out*rate, i.e.
// off' = (1-rate)*off + rate*in, a first-order linear recurrence. Three steps:
//   1. each chunk of CH samples computes b = contribution of its inputs to off at the chunk end;
//   2. one CTA scans the chunk summaries (off_end = a^CH * off_start + b);
//   3. each chunk replays the recurrence from its true starting offset and writes the output.
// ---------------------------------------------------------------------------------------------
constexpr int kDcChunk = 256;
int dc_block_chunks(int count) { return ceil_div(count, kDcChunk); }

__global__ void __launch_bounds__(256)
dc_summary_kernel(const float2* __restrict__ in, int count, float rate, float2* __restrict__ summ) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int n0 = c * kDcChunk;
    if (n0 >= count) return;
    const int n1 = min(count, n0 + kDcChunk);
    const float a = 1.0f - rate;
    float2 b = make_float2(0.0f, 0.0f);
    for (int n = n0; n < n1; n++) {
        const float2 x = in[n];
        b.x = fmaf(a, b.x, rate * x.x);
        b.y = fmaf(a, b.y, rate * x.y);
    }
    summ[c] = b;
}

__global__ void __launch_bounds__(1024)
dc_scan_kernel(float2* __restrict__ summ, int nchunks, int count, float rate, float2* __restrict__ state) {
    // sequential over chunks in one thread per component pair would be nchunks long (<= 3907 for 1e6
    // samples); do it with one warp-free serial loop per CTA thread 0 -- it is ~4k FMAs.
    if (threadIdx.x != 0) return;
    const float a = 1.0f - rate;
    const float aCH = powf(a, (float)kDcChunk);
    float2 off = *state;
    for (int c = 0; c < nchunks; c++) {
        const float2 b = summ[c];
        summ[c] = off; // chunk start offset
        const int len = min(kDcChunk, count - c * kDcChunk);
        const float ap = (len == kDcChunk) ? aCH : powf(a, (float)len);
        off.x = fmaf(ap, off.x, b.x);
        off.y = fmaf(ap, off.y, b.y);
    }
    *state = off;
}

__global__ void __launch_bounds__(256)
dc_apply_kernel(const float2* __restrict__ in, int count, float rate, const float2* __restrict__ summ,
                RingRef dst, uint32_t pos, bool conj) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int n0 = c * kDcChunk;
    if (n0 >= count) return;
    const int n1 = min(count, n0 + kDcChunk);
    float2 off = summ[c];
    for (int n = n0; n < n1; n++) {
        const float2 x = in[n];
        const float2 o = make_float2(x.x - off.x, x.y - off.y);
        ring_store(dst, pos + (uint32_t)n, o, conj);
        off.x = fmaf(o.x, rate, off.x);
        off.y = fmaf(o.y, rate, off.y);
    }
}

cudaError_t launch_dc_block(const float2* in, int count, float rate, float2* state, float2* scratch,
                            RingRef dst, uint32_t pos, bool conj, cudaStream_t st, long long* launches) {
    if (count <= 0) return cudaSuccess;
    const int nch = dc_block_chunks(count);
    dc_summary_kernel<<<ceil_div(nch, 256), 256, 0, st>>>(in, count, rate, scratch);
    dc_scan_kernel<<<1, 32, 0, st>>>(scratch, nch, count, rate, state);
    dc_apply_kernel<<<ceil_div(nch, 256), 256, 0, st>>>(in, count, rate, scratch, dst, pos, conj);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

} // namespace sdrpp
