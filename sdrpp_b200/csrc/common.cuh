// Shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace sdrpp {

// Last error text of the calling thread (returned by sdrpp_cuda_last_error()).
void set_last_error(const std::string& msg);

#define SDRPP_CUDA_TRY(expr)                                                                     \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            ::sdrpp::set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));         \
            return SDRPP_ERR_CUDA;                                                               \
        }                                                                                        \
    } while (0)

__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Opt a kernel in to `bytes` of dynamic shared memory on the CURRENT device. The attribute is per device and the library
// lets every thread pick its own (sdrpp_cuda_init), so what has been set is remembered per (kernel, device), under a lock.
// carveout: cudaFuncAttributePreferredSharedMemoryCarveout for the kernel (percent of the largest setting, or
// cudaSharedmemCarveoutDefault = -1 for the driver's own choice from the kernel's occupancy).
cudaError_t ensure_dynamic_smem(const void* func, size_t bytes, int carveout = 100 /* cudaSharedmemCarveoutMaxShared */);

// Bit-exact sample conversion. Each formula keeps the reference's operation
// order with IEEE round-to-nearest intrinsics so the compiler can neither contract to FMA nor
// replace the divide by a reciprocal multiply (App. C.1).
__device__ __forceinline__ float cvt_u8_rtl(unsigned int v) {
    return __fdiv_rn(__fadd_rn((float)((int)v - 128), 0.5f), 127.5f);
}
__device__ __forceinline__ float cvt_u8_tcp(unsigned int v) {
    return (float)__ddiv_rn((double)v - 128.0, 128.0);
}
__device__ __forceinline__ float cvt_i8(int v) { return __fdiv_rn((float)v, 128.0f); }
__device__ __forceinline__ float cvt_i16_file(int v) {
    return __fdiv_rn(__fadd_rn((float)v, 0.5f), 32767.5f);
}
__device__ __forceinline__ float cvt_i16_volk(int v) { return __fdiv_rn((float)v, 32768.0f); }
__device__ __forceinline__ float cvt_i24_file(int v) { return __fdiv_rn(__fadd_rn((float)v, 0.5f), 8388607.5f); }
__device__ __forceinline__ float cvt_i32_file(int v) { return (float)__ddiv_rn(__dadd_rn((double)v, 0.5), 2147483647.5); }

// Load one complex sample of input format FMT at sample index i and convert it.
// FMT 9 / 10 (internal): SDR++ server wire packets, int8 / int16 divided by a per-packet fp32
// divisor (sample_stream_decompressor.h:23-32; generic VOLK: (float)x / scalar).
template <int FMT>
__device__ __forceinline__ float2 load_sample(const void* __restrict__ base, size_t i, float scale = 1.0f) {
    if constexpr (FMT == 9) {
        const char2 v = __ldg(reinterpret_cast<const char2*>(base) + i);
        return make_float2(__fdiv_rn((float)v.x, scale), __fdiv_rn((float)v.y, scale));
    } else if constexpr (FMT == 10) {
        const short2 v = __ldg(reinterpret_cast<const short2*>(base) + i);
        return make_float2(__fdiv_rn((float)v.x, scale), __fdiv_rn((float)v.y, scale));
    } else if constexpr (FMT == 0) {
        return __ldg(reinterpret_cast<const float2*>(base) + i);
    } else if constexpr (FMT == 1 || FMT == 2) {
        const uchar2 v = __ldg(reinterpret_cast<const uchar2*>(base) + i);
        return FMT == 1 ? make_float2(cvt_u8_rtl(v.x), cvt_u8_rtl(v.y)) : make_float2(cvt_u8_tcp(v.x), cvt_u8_tcp(v.y));
    } else if constexpr (FMT == 3) {
        const char2 v = __ldg(reinterpret_cast<const char2*>(base) + i);
        return make_float2(cvt_i8(v.x), cvt_i8(v.y));
    } else if constexpr (FMT == 4 || FMT == 5) {
        const short2 v = __ldg(reinterpret_cast<const short2*>(base) + i);
        return FMT == 4 ? make_float2(cvt_i16_file(v.x), cvt_i16_file(v.y)) : make_float2(cvt_i16_volk(v.x), cvt_i16_volk(v.y));
    } else if constexpr (FMT == 6) {
        // packed little-endian 24-bit, sign-extended like ((b0 | b1<<8 | b2<<16) << 8) >> 8
        const unsigned char* b = reinterpret_cast<const unsigned char*>(base) + 6 * i;
        const int re = ((int)((unsigned)__ldg(b) | ((unsigned)__ldg(b + 1) << 8) | ((unsigned)__ldg(b + 2) << 16)) << 8) >> 8;
        const int im = ((int)((unsigned)__ldg(b + 3) | ((unsigned)__ldg(b + 4) << 8) | ((unsigned)__ldg(b + 5) << 16)) << 8) >> 8;
        return make_float2(cvt_i24_file(re), cvt_i24_file(im));
    } else if constexpr (FMT == 7) {
        const int2 v = __ldg(reinterpret_cast<const int2*>(base) + i);
        return make_float2(cvt_i32_file(v.x), cvt_i32_file(v.y));
    } else {
        const double2 v = __ldg(reinterpret_cast<const double2*>(base) + i);
        return make_float2((float)v.x, (float)v.y);
    }
}

__host__ __device__ constexpr int fmt_bytes_per_sample(int fmt) {
    return fmt == 0 ? 8 : (fmt == 1 || fmt == 2 || fmt == 3 || fmt == 9) ? 2 : (fmt == 4 || fmt == 5 || fmt == 10) ? 4 : fmt == 6 ? 6 : fmt == 7 ? 8 : 16;
}

constexpr int kFmtPcmI8 = 9, kFmtPcmI16 = 10; // internal formats behind sdrpp_cuda_pcm_* (run-time divisor)

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

} // namespace sdrpp
