// Channelizer stage 1 on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Replaces, for every VFO of a group at once, FrequencyXlator::process (dsp/channel/frequency_xlator.h:43-50)
// followed by the first DecimatingFIR of the PowerDecimator cascade (dsp/filter/decimating_fir.h:45-68,
// dsp/multirate/power_decimator.h:51-67), like stage1_kernel in channelizer.cu -- which stays as the path for
// plan shapes and blocks this one does not take (small first-stage decimation, windows that reach before a
// VFO's epoch). Same results contract (<= 1e-5 relative RMS against the reference's fp32 blocks).
//
// The sum as a matrix product. Cut the input into rows of D samples aligned to absolute multiples of D.
// An output's window starts s = n0 mod D samples into a row, so with the taps shifted by s (h'[k] = h[k-s])
// and k = a*D + p (a < A = ceil((T+s)/D), p < D):
//     y[m] = e^{j phi(row R_m)} * sum_a V[R_m + a][a],
//     V[R][a] = sum_p x[R*D + p] * ( h'[a*D+p] * e^{j w (a*D+p)} )            (complex * complex)
// V = X * B with X[R][p] the untranslated samples (shared by EVERY VFO and plan of the same D) and
// B[p][(vfo, a)] a per-VFO constant table: a complex GEMM with M = rows, K = D, N = A * VFOs, written as a
// real one with K = 2D (re, im interleaved exactly as the samples lie in memory) and N = 2*A per VFO.
// fp32 accuracy on fp16 tensor cores: both operands are split into two fp16 halves (hi + lo, 22 significant
// bits, block-scaled by a power of two per 8 rows of X and per group for B) and three products are
// accumulated in fp32: Xhi*Bhi + Xhi*Blo + Xlo*Bhi (the dropped Xlo*Blo term is 2^-22 relative).
//
//  s1t_split_kernel   cf32 ring -> fp16 hi/lo planes stored directly in the UMMA K-major SWIZZLE_128B
//                     shared-memory image (an A tile of 128 rows x 64 K is one contiguous 16 KB bulk copy)
//  s1t_build_b_kernel per VFO tile of 16 VFOs: the B image (hi/lo, shifted taps x phasors), rebuilt on retune
//  s1t_kernel         persistent, one CTA per SM, warp-specialised: bulk-copy producer | MMA issuer |
//                     8 epilogue warps (TMEM -> registers, sum over a across rows by warp shuffles,
//                     rotate by the row phase, store to the VFO slabs); TMEM accumulator double-buffered.
#include "common.cuh"
#include "kernels.h"
#include <cuda_fp16.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

namespace sdrpp {

namespace {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// A wait that never completes traps (the launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++)
        if (spins > (1u << 22)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], fp16 operands, fp32 accumulate; issued by ONE thread for the CTA
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row atoms of 1024 B stacked every 1024 B
// (cute/arch/mma_sm100_desc.hpp SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1
// [46,48), layout type [61,64) with SWIZZLE_128B = 2).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// Instruction descriptor (InstrDescriptor): D = F32 [4,6), A/B = F16 (0), K-major both, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float2 phasor64(uint64_t phase) {
    const float x = (float)(int32_t)(phase >> 32) * 4.656612873077393e-10f; // turns*2^64 -> half turns in [-1, 1)
    float s, c;
    sincospif(x, &s, &c);
    return make_float2(c, s);
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

#ifdef SDRPP_S1T_TRACE
// Debug build only (-DSDRPP_S1T_TRACE): per-CTA cycle counters of the three roles (tools/s1t_trace.py)
__device__ long long g_s1t_trace[256][24];
#define S1T_T0(var) const long long var = clock64()
#define S1T_ACC(slot, var) do { s1t_acc_[slot] += clock64() - (var); } while (0)
#define S1T_DECL long long s1t_acc_[24] = { 0 }
#define S1T_ARG , s1t_acc_
#define S1T_FLUSH(slot) do { if ((threadIdx.x & 31) == 0 && blockIdx.x < 256) g_s1t_trace[blockIdx.x][slot] = s1t_acc_[slot]; } while (0)
#define S1T_SET(slot, val) do { if (blockIdx.x < 256) g_s1t_trace[blockIdx.x][slot] = (val); } while (0)
#else
#define S1T_T0(var) do {} while (0)
#define S1T_ACC(slot, var) do {} while (0)
#define S1T_DECL do {} while (0)
#define S1T_ARG
#define S1T_FLUSH(slot) do {} while (0)
#define S1T_SET(slot, val) do {} while (0)
#endif

constexpr int kRowsPerTile = 128;   // MMA M
constexpr int kOutPerTile = 120;    // outputs per time tile: 128 - (A-1) rounded down to whole 8-row atoms
constexpr int kNV = kS1TVfosPerTile;
constexpr int kChunkBytes = 16384;  // one A-operand chunk: 128 rows x 64 fp16
constexpr int kMaxChunks = 8;
constexpr int kEpiWarps = 16;       // 4 per TMEM lane quadrant, each with kNV/4 VFOs of the tile
constexpr int kThreads = 64 + 32 * kEpiWarps; // warp 0 producer, warp 1 MMA issuer, warps 2.. epilogue
constexpr int kNVW = kNV / (kEpiWarps / 4); // VFOs per epilogue warp
constexpr int kEpiBatch = 2;         // VFOs per TMEM load batch in the epilogue
constexpr int kXchFloats = 2 /*buffers*/ * 4 /*quadrants*/ * 7 /*lanes*/ * kNV * 2;

} // namespace

// ---------------------------------------------------------------------------------------------
// cf32 ring -> fp16 hi/lo planes in the UMMA shared-memory image.
// One CTA per 8-row group (8*D samples): block scale 2^e from the group's largest component, then
// hi = fp16(x*2^e), lo = fp16(x*2^e - hi). A thread converts 4 consecutive samples = one 16-byte chunk
// (8 fp16: re0 im0 .. re3 im3) of the row and stores it at the swizzled chunk position (chunk ^ row).
// Samples at or beyond abs_end (not written yet) are stored as zeros.
// ---------------------------------------------------------------------------------------------
struct S1TSplitArgs {    // per-block arguments, read through the launcher's descriptor
    RingRef ring;
    S1TPlanes pl;
    int64_t g_first, abs_end;
    // fused ingest (cf32 blocks that need no conversion): samples at or after abs_block come from the block itself and are
    // written to the ring on the way -- the separate 4.9 MB ring copy and its launch are gone
    const float2* raw;
    int64_t abs_block;
    int conj;
};

template <bool FUSED>
__global__ void __launch_bounds__(256)
s1t_split_kernel(const S1TSplitArgs* __restrict__ ap) {
    __shared__ float red[8];
    const RingRef ring = ap->ring;
    const S1TPlanes pl = ap->pl;
    const int64_t g_first = ap->g_first, abs_end = ap->abs_end;
    const int D = pl.D;
    const int t = threadIdx.x;
    const int64_t g = g_first + blockIdx.x;
    const int cpr = D >> 2;               // 16-byte chunks per row and plane
    const int r = t / cpr, c16 = t - r * cpr;
    const int64_t n0 = g * (int64_t)(8 * D) + (int64_t)r * D + (int64_t)c16 * 4 + pl.origin;
    float v[8];
    const uint32_t i0 = (uint32_t)((uint64_t)n0 & ring.mask);
    bool from_ring = true;
    if constexpr (FUSED) {
        const int64_t abs_block = ap->abs_block;
        if (n0 + 3 >= abs_block && n0 < abs_end) {
            from_ring = false;
            const float2* __restrict__ raw = ap->raw;
            const float sgn = ap->conj ? -1.0f : 1.0f;
            if (n0 >= abs_block && n0 + 3 < abs_end) {
                const int64_t k = n0 - abs_block;
                if ((k & 1) == 0) {
                    const float4 a = __ldg(reinterpret_cast<const float4*>(raw + k));
                    const float4 b = __ldg(reinterpret_cast<const float4*>(raw + k + 2));
                    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
                } else {
#pragma unroll
                    for (int i = 0; i < 4; i++) { const float2 x = __ldg(raw + k + i); v[2 * i] = x.x; v[2 * i + 1] = x.y; }
                }
#pragma unroll
                for (int i = 0; i < 4; i++) v[2 * i + 1] *= sgn;
                // n0 is a multiple of 4 and so is the ring length: 16-byte aligned, never wraps inside the quad
                float4* o = reinterpret_cast<float4*>(ring.base + i0);
                o[0] = make_float4(v[0], v[1], v[2], v[3]);
                o[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else {
                // the quad straddles the start or the end of the block
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int64_t n = n0 + i;
                    float2 x = make_float2(0.0f, 0.0f);
                    if (n >= abs_block && n < abs_end) {
                        x = __ldg(raw + (n - abs_block));
                        x.y *= sgn;
                        ring.base[(i0 + (uint32_t)i) & ring.mask] = x;
                    } else if (n >= 0 && n < abs_block) {
                        x = ring.base[(i0 + (uint32_t)i) & ring.mask];
                    }
                    v[2 * i] = x.x; v[2 * i + 1] = x.y;
                }
            }
        }
    }
    if (from_ring) {
        const float4 a = *reinterpret_cast<const float4*>(ring.base + i0);
        const float4 b = *reinterpret_cast<const float4*>(ring.base + i0 + 2);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    float m = 0.0f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (n0 + i >= abs_end || n0 + i < 0) { v[2 * i] = 0.0f; v[2 * i + 1] = 0.0f; }
        m = fmaxf(m, fmaxf(fabsf(v[2 * i]), fabsf(v[2 * i + 1])));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((t & 31) == 0) red[t >> 5] = m;
    __syncthreads();
    const int nw = (int)blockDim.x >> 5;
    m = red[0];
    for (int w = 1; w < nw; w++) m = fmaxf(m, red[w]);
    // largest component scaled into [2^13, 2^14): fp16 keeps 11 bits down to 2^-14 and loses range above 2^16
    int e = 0;
    if (m > 0.0f && m < 3.0e38f) {
        const int ex = (int)((__float_as_uint(m) >> 23) & 0xffu) - 126; // m = f * 2^ex, f in [0.5, 1)
        e = 14 - ex;
        e = e > 60 ? 60 : (e < -40 ? -40 : e);
    }
    const float sc = __uint_as_float((uint32_t)(127 + e) << 23);
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float a = v[2 * i] * sc, b = v[2 * i + 1] * sc;
        const __half2 h = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(h);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h);
        lo[i] = pack_half2(a - hf.x, b - hf.y);
    }
    const uint32_t slot = (uint32_t)((uint64_t)g & pl.group_mask);
    const int kh = c16 >> 3, ch = c16 & 7;
    const size_t off = ((size_t)kh * ((size_t)pl.group_mask + 1) + slot) * 1024 + (size_t)r * 128 + (size_t)((ch ^ r) << 4);
    *reinterpret_cast<uint4*>(pl.hi + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(pl.lo + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    if (t == 0) pl.sinv[slot] = __uint_as_float((uint32_t)(127 - e) << 23);
}

// raw != nullptr: the block [abs_block, abs_end) is still in the caller's cf32 buffer `raw` and this launch also writes it
// to the ring (the caller has checked with s1t_split_covers that every new sample lies in a converted group)
cudaError_t launch_s1t_split(Launcher& L, int sid, RingRef ring, const S1TPlanes& pl, int64_t abs_begin, int64_t abs_end,
                             const void* raw, int64_t abs_block, bool conj) {
    if (abs_end <= abs_begin) return cudaSuccess;
    const int gs = 8 * pl.D;
    if (abs_end <= pl.origin) return cudaSuccess;
    const int64_t g0 = std::max<int64_t>(abs_begin - pl.origin, 0) / gs, g1 = (abs_end - 1 - pl.origin) / gs;
    const S1TSplitArgs* d = L.push(S1TSplitArgs{ ring, pl, g0, abs_end, reinterpret_cast<const float2*>(raw), abs_block, conj ? 1 : 0 });
    if (!d) return cudaErrorMemoryAllocation;
    const void* fn = raw ? (const void*)s1t_split_kernel<true> : (const void*)s1t_split_kernel<false>;
    return L.kernel(sid, fn, dim3((unsigned)(g1 - g0 + 1)), dim3((unsigned)(2 * pl.D)), 0, d);
}

// Does a split launch over [abs_begin, abs_end) convert every sample from abs_block on? (Not at the very start of a stream,
// where the first group begins at the row origin.)
bool s1t_split_covers(const S1TPlanes& pl, int64_t abs_begin, int64_t abs_block, int64_t abs_end) {
    if (abs_end <= abs_begin || abs_end <= pl.origin || abs_begin > abs_block) return false;
    const int gs = 8 * pl.D;
    const int64_t g0 = std::max<int64_t>(abs_begin - pl.origin, 0) / gs;
    return g0 * gs + pl.origin <= abs_block;
}

// ---------------------------------------------------------------------------------------------
// B image of one group: for every tile of 16 VFOs, planes (hi, lo) x k-halves x N rows of 128 bytes,
// row n = (vfo_local, a, c) with c = 0 the real and c = 1 the imaginary output column, K index 2p + ci
// multiplying (ci = 0) x.re or (ci = 1) x.im of sample p of the row:
//     w = h'[a*D+p] * e^{j*dphi*(a*D+p)}:   c=0: (+w.re, -w.im)    c=1: (+w.im, +w.re)
// One thread per 16-byte chunk (4 samples x 2).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
s1t_build_b_kernel(uint8_t* blob, const VfoDev* vfos, int nvfo, const float* taps, int T, int D, int shift, int A, int escale) {
    const int N = 2 * A * kNV, NKH = D >> 5;
    const int chunks_per_tile = NKH * N * 8;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int vt = blockIdx.y;
    if (idx >= chunks_per_tile) return;
    const int ch = idx & 7, n = (idx >> 3) % N, kh = idx / (8 * N);
    const int vl = n / (2 * A), a = (n - vl * 2 * A) >> 1, c = n & 1;
    const int v = vt * kNV + vl;
    const float sc = __uint_as_float((uint32_t)(127 + escale) << 23);
    float val[8];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int p = kh * 32 + ch * 4 + i;
        const int k = a * D + p, kt = k - shift;
        float wr = 0.0f, wi = 0.0f;
        if (v < nvfo && kt >= 0 && kt < T) {
            const float2 e = phasor64((uint64_t)k * vfos[v].dphi);
            const float h = taps[kt] * sc;
            wr = h * e.x; wi = h * e.y;
        }
        val[2 * i] = c == 0 ? wr : wi;
        val[2 * i + 1] = c == 0 ? -wi : wr;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const __half2 h = __floats2half2_rn(val[2 * i], val[2 * i + 1]);
        const float2 hf = __half22float2(h);
        hi[i] = *reinterpret_cast<const uint32_t*>(&h);
        lo[i] = pack_half2(val[2 * i] - hf.x, val[2 * i + 1] - hf.y);
    }
    const size_t plane = (size_t)NKH * N * 128;
    uint8_t* tile = blob + (size_t)vt * 2 * plane;
    const size_t off = ((size_t)kh * N + n) * 128 + (size_t)((ch ^ (n & 7)) << 4);
    *reinterpret_cast<uint4*>(tile + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(tile + plane + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

size_t s1t_b_bytes(int A, int D, int nvfo) {
    return (size_t)ceil_div(nvfo, kNV) * 2 * (size_t)(D >> 5) * (size_t)(2 * A * kNV) * 128;
}
bool s1t_supported(int T, int D) { return (D == 32 || D == 64) && ceil_div(T + D - 1, D) <= 8 && T >= D; }
int s1t_A(int T, int D, int shift) { return ceil_div(T + shift, D); }
// exponent that puts the largest tap into [2^13, 2^14) (the phasor has unit modulus)
int s1t_b_exponent(const float* taps, int T) {
    float m = 0.0f;
    for (int i = 0; i < T; i++) m = std::max(m, std::fabs(taps[i]));
    if (!(m > 0.0f)) return 0;
    int ex = 0;
    std::frexp(m, &ex);
    return std::max(-40, std::min(60, 14 - ex));
}

cudaError_t launch_s1t_build_b(Launcher& L, int sid, uint8_t* blob, const VfoDev* vfos, int nvfo, const float* d_taps, int T, int D, int shift,
                               int A, int escale) {
    const int N = 2 * A * kNV, NKH = D >> 5;
    dim3 grid(ceil_div(NKH * N * 8, 256), ceil_div(nvfo, kNV));
    return L.kernel(sid, (const void*)s1t_build_b_kernel, grid, dim3(256), 0, blob, vfos, nvfo, d_taps, T, D, shift, A, escale);
}

// ---------------------------------------------------------------------------------------------
// Epilogue of one time tile for one epilogue warp: quadrant q (TMEM lanes 32q..32q+31 = tile rows), part hf
// (VFOs kNVW*hf .. of the tile). Thread = one row R. Per VFO: load the 2A accumulator columns and sum
// V[R+a][a] over a with warp shuffles, each term weighted by the block scale of the row it comes from; the part
// that reaches into the next warp's rows is handed over through shared memory (written by the lanes it wraps
// onto). cur[j] = e^{j phi(R)} of the thread's row for VFO j, advanced by the caller from tile to tile; sc = block scale
// of the thread's row (loaded one tile ahead by the caller), m = output index of the row, obase[j] = where output 0 of
// VFO j goes (null: no such VFO). Everything the tile needs besides the accumulator sits in registers: a shared-memory
// or constant-bank round trip here queues behind the MMA operand stream (tools/s1t_trace.py: the store phase took 785
// cycles per tile while it fetched its pointers from shared memory, the scale load 290).
// ---------------------------------------------------------------------------------------------
template <int A>
__device__ __forceinline__ void s1t_epilogue_tile(uint32_t tmem_acc, int q, int hf, int lane, float sc, int m, int M, float* xch,
                                                  uint64_t* tempty_bar, const float2* cur, float2* const* obase
#ifdef SDRPP_S1T_TRACE
                                                  , long long* s1t_acc_
#endif
                                                  ) {
    constexpr int NVH = kNVW;
#ifdef SDRPP_S1T_TRACE
    const long long tl0_ = clock64();
#endif
    const int row = q * 32 + lane;
    // weight of the term taken from lane (lane + a) & 31: its row's scale, routed to the in-warp sum (scA) or to
    // the sum that belongs to the previous quadrant's row (scW)
    float scA[A], scW[A];
#pragma unroll
    for (int a = 0; a < A; a++) {
        const float s = a ? __shfl_sync(0xffffffffu, sc, (lane + a) & 31) : sc;
        const bool wrapped = lane + a >= 32;
        scA[a] = wrapped ? 0.0f : s;
        scW[a] = wrapped ? s : 0.0f;
    }
#ifdef SDRPP_S1T_TRACE
    if (q == 2 && hf == 0 && lane == 0) { s1t_acc_[16] += (long long)(scA[0] != 12345.0f) * (clock64() - tl0_); }
#endif
    const uint32_t t0 = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(hf * NVH * 2 * A);
    float2 y[NVH];
    float2* xw = reinterpret_cast<float2*>(xch) + ((hf * 4 + q) * 7) * NVH;
#pragma unroll
    for (int vb = 0; vb < NVH; vb += kEpiBatch) {
        uint32_t r[kEpiBatch][16];
#pragma unroll
        for (int j = 0; j < kEpiBatch; j++) tc_ld16(t0 + (uint32_t)((vb + j) * 2 * A), r[j]);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < kEpiBatch; j++) {
            float sre = __uint_as_float(r[j][0]) * scA[0], sim = __uint_as_float(r[j][1]) * scA[0];
            float wre = 0.0f, wim = 0.0f;
#pragma unroll
            for (int a = 1; a < A; a++) {
                const float tre = __shfl_sync(0xffffffffu, __uint_as_float(r[j][2 * a]), (lane + a) & 31);
                const float tim = __shfl_sync(0xffffffffu, __uint_as_float(r[j][2 * a + 1]), (lane + a) & 31);
                sre = fmaf(tre, scA[a], sre); sim = fmaf(tim, scA[a], sim);
                wre = fmaf(tre, scW[a], wre); wim = fmaf(tim, scW[a], wim);
            }
            y[vb + j] = make_float2(sre, sim);
            // lanes 25..31 hold the sums that belong to rows 25..31 of the PREVIOUS quadrant
            if (lane >= 25) xw[(lane - 25) * NVH + vb + j] = make_float2(wre, wim);
        }
    }
    // accumulator stage drained: hand it back to the MMA issuer
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty_bar);
#ifdef SDRPP_S1T_TRACE
    const long long tb_ = clock64();
    if (q == 2 && hf == 0 && lane == 0) { S1T_ACC(6, tl0_); }
#endif
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory"); // the epilogue warps: wrapped sums visible
#ifdef SDRPP_S1T_TRACE
    if (q == 2 && hf == 0 && lane == 0) { S1T_ACC(7, tb_); }
    const long long ts_ = clock64();
#endif
    const bool out_row = row < kOutPerTile && m >= 0 && m < M;
    const bool take = q < 3 && lane >= 25;
    const float2* xr = reinterpret_cast<const float2*>(xch) + ((hf * 4 + q + 1) * 7) * NVH; // next quadrant's hand-over
#pragma unroll
    for (int j = 0; j < NVH; j++) {
        float2 t = y[j];
        if (take) {
            const float2 w = xr[(lane - 25) * NVH + j];
            t.x += w.x; t.y += w.y;
        }
        float2* o = obase[j];
        if (out_row && o) {
            const float2 e = cur[j];
            __stcg(o + m, make_float2(t.x * e.x - t.y * e.y, t.x * e.y + t.y * e.x));
        }
    }
#ifdef SDRPP_S1T_TRACE
    if (q == 2 && hf == 0 && lane == 0) { S1T_ACC(17, ts_); }
#endif
}

// ---------------------------------------------------------------------------------------------
// Persistent warp-specialised kernel. CTA b serves one VFO tile (its B image stays in shared memory) of one
// group over a contiguous range of time tiles.
// ---------------------------------------------------------------------------------------------
template <int NKH>
// Experiment knob (profiles/r2m_coresidency.txt): -DSDRPP_S1T_MAXNREG=80 (the compiler's own choice is 96; 16-20 bytes of
// spills, same kernel time) together with SDRPP_S1T_SMEM_CAP=190000 leaves 19 K registers and 42 KB of shared memory per SM
// for a narrow tail CTA of the previous block.
#ifdef SDRPP_S1T_MAXNREG
__global__ void __maxnreg__(SDRPP_S1T_MAXNREG)
#else
__global__ void __launch_bounds__(kThreads, 1)
#endif
s1t_kernel(const S1TArgs* __restrict__ ap) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const S1TArgs& a = *ap;   // per-block arguments in the launcher's descriptor (device memory): read once per role
    int gi = 0;
    while (gi + 1 < a.ngroups && (int)blockIdx.x >= a.g[gi + 1].cta_begin) gi++;
    const S1TGroupArgs& G = a.g[gi];
    const int local = (int)blockIdx.x - G.cta_begin;
    const int vt = local / G.cta_per_vtile, ts = local - vt * G.cta_per_vtile;
    const int tt0 = (int)((int64_t)G.n_ttiles * ts / G.cta_per_vtile);
    const int tt1 = (int)((int64_t)G.n_ttiles * (ts + 1) / G.cta_per_vtile);
    const int A = G.A, N = 2 * A * kNV;
    const int nch = a.nchunks;
    const uint32_t b_plane = (uint32_t)NKH * (uint32_t)N * 128u;   // bytes of one B plane (hi or lo)

    // the dynamic shared memory base is 1024-byte aligned (SWIZZLE_128B atoms); pointers derived from smem_raw keep
    // the shared address space, so the epilogue's accesses compile to LDS/STS with 32-bit addressing
    if ((smem_addr(smem_raw) & 1023u) != 0u) __trap();
    uint8_t* smB = smem_raw;                              // [hi|lo][kh][N][128]
    uint8_t* smA = smB + 2 * b_plane;                     // nch chunks of 16 KB (multiple of 1024: N % 8 == 0)
    float* xch = reinterpret_cast<float*>(smA + (size_t)nch * kChunkBytes);
    ulonglong2* nco = reinterpret_cast<ulonglong2*>(xch + kXchFloats);   // [kNV] (P0, W) of the tile's VFOs
    uint64_t* bars = reinterpret_cast<uint64_t*>(nco + kNV);
    uint64_t* full = bars;                 // [kMaxChunks]
    uint64_t* empty = bars + kMaxChunks;   // [kMaxChunks]
    uint64_t* bfull = bars + 2 * kMaxChunks;
    uint64_t* tfull = bfull + 1;           // [2]
    uint64_t* tempty = tfull + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    S1T_DECL;
#ifdef SDRPP_S1T_TRACE
    const long long tk0_ = clock64();
    if (tid == 0 && blockIdx.x < 256) {
        for (int i = 0; i < 24; i++) g_s1t_trace[blockIdx.x][i] = 0;
        g_s1t_trace[blockIdx.x][12] = tt1 - tt0; g_s1t_trace[blockIdx.x][13] = A;
        long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));
        g_s1t_trace[blockIdx.x][11] = gt_;
    }
    __syncthreads();
#endif
    if (tid == 0) {
        for (int i = 0; i < kMaxChunks; i++) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        mbar_init(bfull, 1);
        for (int i = 0; i < 2; i++) { mbar_init(tfull + i, 1); mbar_init(tempty + i, kEpiWarps); }
        mbar_fence_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // The producer and MMA warps run their loops convergently (all 32 lanes wait on the barriers) and elect one lane
    // only for the issuing instructions: bulk copies and tcgen05.mma take their operands from uniform registers, and
    // in divergent code every one of them costs a register-to-uniform broadcast loop.
    if (warp == 0) {
        // ===== producer: B image once, then the A chunks of every time tile (hi, lo per k-half) =====
        if (tt1 > tt0) {
            const uint8_t* bsrc = G.bblob + (size_t)vt * 2 * b_plane;
            const uint32_t piece = (uint32_t)N * 128u;
            if (elect_one()) {
                mbar_expect_tx(bfull, 2 * b_plane);
                for (int i = 0; i < 2 * NKH; i++) bulk_g2s(smB + (size_t)i * piece, bsrc + (size_t)i * piece, piece, bfull);
            }
            __syncwarp();
            // the arguments live in device memory now: keep what the loop needs in registers (every barrier wait below is a
            // compiler memory fence, so a field read inside the loop would be loaded again on every pass)
            const uint32_t gmask = a.pl.group_mask, ng = gmask + 1;
            const uint8_t* const pl_hi = a.pl.hi;
            const uint8_t* const pl_lo = a.pl.lo;
            const int64_t g_row0 = G.row0;
            uint32_t slot = 0, ph = 0; // ring slot and its phase bit, advanced without divisions
            for (int tt = tt0; tt < tt1; tt++) {
                const int64_t row_t = g_row0 + (int64_t)kOutPerTile * tt;
                const uint32_t gs = (uint32_t)((uint64_t)(row_t >> 3) & gmask);
                const uint32_t n1 = min(16u, ng - gs);   // groups before the ring wraps
#pragma unroll 1
                for (int c = 0; c < 2 * NKH; c++) {
                    S1T_T0(tw_);
                    mbar_wait(empty + slot, ph ^ 1u);
                    S1T_ACC(10, tw_);
                    const uint8_t* plane = ((c & 1) ? pl_lo : pl_hi) + (size_t)(c >> 1) * ng * 1024;
                    uint8_t* dst = smA + (size_t)slot * kChunkBytes;
                    if (elect_one()) {
                        mbar_expect_tx(full + slot, (uint32_t)kChunkBytes);
                        bulk_g2s(dst, plane + (size_t)gs * 1024, n1 * 1024u, full + slot);
                        if (n1 < 16u) bulk_g2s(dst + (size_t)n1 * 1024, plane, (16u - n1) * 1024u, full + slot);
                    }
                    __syncwarp();
                    if (++slot == (uint32_t)nch) { slot = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: Xhi*Bhi + Xhi*Blo + Xlo*Bhi per k-half, fp32 accumulate in TMEM =====
        if (tt1 > tt0) {
            const uint32_t idesc = umma_idesc_f16(kRowsPerTile, N);
            mbar_wait(bfull, 0);
#ifdef SDRPP_S1T_TRACE
            if (lane == 0) { S1T_ACC(14, tk0_); }
#endif
            tc_fence_after();
            const uint32_t sB = smem_addr(smB), sA = smem_addr(smA);
            uint32_t slot = 0, ph = 0;
            for (int tt = tt0; tt < tt1; tt++) {
                const uint32_t k = (uint32_t)(tt - tt0), as = k & 1u, aph = (k >> 1) & 1u;
                S1T_T0(tw_);
                mbar_wait(tempty + as, aph ^ 1u);
                S1T_ACC(2, tw_);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256u;
#pragma unroll 1
                for (int kh = 0; kh < NKH; kh++) {
                    const uint64_t bh = umma_desc_sw128(sB + (uint32_t)kh * (uint32_t)N * 128u);
                    const uint64_t bl = umma_desc_sw128(sB + b_plane + (uint32_t)kh * (uint32_t)N * 128u);
                    {
                        S1T_T0(tf_);
                        mbar_wait(full + slot, ph);
                        S1T_ACC(3, tf_);
                        tc_fence_after();
                        const uint64_t xh = umma_desc_sw128(sA + slot * (uint32_t)kChunkBytes);
                        S1T_T0(ti_);
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < 4; ks++) {
                                tc_mma_f16(d_tmem, xh + 2u * ks, bh + 2u * ks, idesc, (kh | ks) ? 1u : 0u);
                                tc_mma_f16(d_tmem, xh + 2u * ks, bl + 2u * ks, idesc, 1u);
                            }
                            tc_commit(empty + slot);
                        }
                        __syncwarp();
                        S1T_ACC(4, ti_);
                        if (++slot == (uint32_t)nch) { slot = 0; ph ^= 1u; }
                    }
                    {
                        S1T_T0(tf_);
                        mbar_wait(full + slot, ph);
                        S1T_ACC(3, tf_);
                        tc_fence_after();
                        const uint64_t xl = umma_desc_sw128(sA + slot * (uint32_t)kChunkBytes);
                        S1T_T0(ti_);
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < 4; ks++) tc_mma_f16(d_tmem, xl + 2u * ks, bh + 2u * ks, idesc, 1u);
                            tc_commit(empty + slot);
                            if (kh == NKH - 1) tc_commit(tfull + as);
                        }
                        __syncwarp();
                        S1T_ACC(8, ti_);
                        if (++slot == (uint32_t)nch) { slot = 0; ph ^= 1u; }
                    }
                }
            }
        }
    } else {
        // ===== epilogue warps =====
        const int q = warp & 3, hf = (warp - 2) >> 2;
        constexpr int NVH = kNVW;
        const int row = q * 32 + lane;
        const int64_t g_row0 = G.row0;
        const int g_M = G.M, g_nvfo = G.nvfo, D = a.pl.D;
        const float b_scale_inv = G.b_scale_inv;
        const VfoDev* g_vfos = G.vfos;
        const float* __restrict__ sinv = a.pl.sinv;
        const uint32_t gmask = a.pl.group_mask;
        // output index of this thread's row in time tile 0 (advances by kOutPerTile per tile); fits 32 bits: |row0 - row_first| < 8
        int m = (int)(g_row0 - G.row_first) + row + kOutPerTile * tt0;
        // NCO phase of this thread's row per VFO: exact (64-bit accumulator) every 8 tiles, advanced by the constant tile
        // step in between. phase(R) = P0 + R * W with P0 = phi_ref + (origin - n_ref) * dphi, W = D * dphi (mod 2^64): the
        // two words per VFO are computed once per CTA and kept in shared memory, so the refresh has no global round trip.
        float2 cur[NVH], stp[NVH];
        float2* obase[NVH];
        if (tid - 64 < kNV) {
            const int v = vt * kNV + (tid - 64);
            uint64_t p0 = 0, w = 0;
            if (v < g_nvfo) {
                const VfoDev* vd = g_vfos + v;
                w = (uint64_t)D * vd->dphi;
                p0 = vd->phi_ref + (uint64_t)((int64_t)a.pl.origin - vd->n_ref) * vd->dphi;
            }
            nco[tid - 64] = make_ulonglong2(p0, w);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
        auto exact_phase = [&](int tt) {
            const uint64_t R = (uint64_t)(g_row0 + (int64_t)kOutPerTile * tt + row);
#pragma unroll
            for (int j = 0; j < NVH; j++) {
                const ulonglong2 c = nco[hf * NVH + j];
                cur[j] = phasor64(c.x + R * c.y);
            }
        };
#pragma unroll
        for (int j = 0; j < NVH; j++) {
            const int v = vt * kNV + hf * NVH + j;
            stp[j] = phasor64((uint64_t)kOutPerTile * nco[hf * NVH + j].y);
            obase[j] = v < g_nvfo ? g_vfos[v].slab + G.out_off : nullptr;
        }
        auto row_scale = [&](int tt) {
            const int64_t R = g_row0 + (int64_t)kOutPerTile * tt + row;
            return __ldg(sinv + (uint32_t)((uint64_t)(R >> 3) & gmask)) * b_scale_inv;
        };
        float sc_next = tt1 > tt0 ? row_scale(tt0) : 0.0f;
        for (int tt = tt0; tt < tt1; tt++) {
            const uint32_t k = (uint32_t)(tt - tt0), as = k & 1u, aph = (k >> 1) & 1u;
#ifdef SDRPP_S1T_TRACE
            const long long tx_ = clock64();
#endif
            if ((k & 7u) == 0u) exact_phase(tt);
#ifdef SDRPP_S1T_TRACE
            if (warp == 2 && lane == 0) { s1t_acc_[18] += (long long)(cur[0].x != 12345.0f) * (clock64() - tx_); }
#endif
            const float sc = sc_next;
            if (tt + 1 < tt1) sc_next = row_scale(tt + 1);   // in flight while this tile is summed
#ifdef SDRPP_S1T_TRACE
            const bool tr_ = warp == 2 && lane == 0;
            const long long te_ = clock64();
#endif
            mbar_wait(tfull + as, aph);
#ifdef SDRPP_S1T_TRACE
            if (tr_ && k == 0) { S1T_ACC(15, tk0_); }
            if (tr_) { S1T_ACC(5, te_); }
            const long long tl_ = clock64();
#endif
            tc_fence_after();
            float* xb = xch + (k & 1u) * (kXchFloats / 2);
            const uint32_t acc = tmem_base + as * 256u;
            switch (A) {
            case 4: s1t_epilogue_tile<4>(acc, q, hf, lane, sc, m, g_M, xb, tempty + as, cur, obase S1T_ARG); break;
            case 5: s1t_epilogue_tile<5>(acc, q, hf, lane, sc, m, g_M, xb, tempty + as, cur, obase S1T_ARG); break;
            case 6: s1t_epilogue_tile<6>(acc, q, hf, lane, sc, m, g_M, xb, tempty + as, cur, obase S1T_ARG); break;
            case 7: s1t_epilogue_tile<7>(acc, q, hf, lane, sc, m, g_M, xb, tempty + as, cur, obase S1T_ARG); break;
            default: s1t_epilogue_tile<8>(acc, q, hf, lane, sc, m, g_M, xb, tempty + as, cur, obase S1T_ARG); break;
            }
            m += kOutPerTile;
#pragma unroll
            for (int j = 0; j < NVH; j++) cur[j] = cmul(cur[j], stp[j]);
#ifdef SDRPP_S1T_TRACE
            if (tr_) { S1T_ACC(9, tl_); }
#endif
        }
    }
#ifdef SDRPP_S1T_TRACE
    if (warp == 0) { S1T_FLUSH(10); }
    if (warp == 1) { S1T_FLUSH(2); S1T_FLUSH(3); S1T_FLUSH(4); S1T_FLUSH(8); S1T_FLUSH(14); }
    if (warp == 2) { S1T_FLUSH(5); S1T_FLUSH(6); S1T_FLUSH(7); S1T_FLUSH(9); S1T_FLUSH(15); S1T_FLUSH(16); S1T_FLUSH(17); S1T_FLUSH(18); }
#endif
    tc_fence_before();
    __syncthreads();
#ifdef SDRPP_S1T_TRACE
    if (tid == 0) {
        S1T_ACC(1, tk0_); S1T_FLUSH(1);
        long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));
        if (blockIdx.x < 256) g_s1t_trace[blockIdx.x][0] = gt_;
    }
#endif
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

static size_t s1t_smem_bytes(int NKH, int A, int nchunks) {
    return 1024 + (size_t)2 * NKH * (2 * A * kNV) * 128 + (size_t)nchunks * kChunkBytes + kXchFloats * sizeof(float) + kNV * sizeof(ulonglong2) + 256;
}

// Fills in cta_begin / cta_per_vtile (time tiles of a VFO tile are split between CTAs so that every SM gets
// about the same number of MMA cycles) and launches one persistent grid for all groups.
cudaError_t launch_s1t(Launcher& L, int sid, S1TArgs& a, int num_sms) {
    if (a.ngroups <= 0) return cudaSuccess;
    const int NKH = a.pl.D >> 5;
    if (NKH != 1 && NKH != 2) return cudaErrorInvalidValue;
    int maxA = 0;
    double total = 0.0;
    for (int i = 0; i < a.ngroups; i++) {
        S1TGroupArgs& g = a.g[i];
        if (g.A < 4 || g.A > 8 || g.n_ttiles <= 0 || g.nvfo <= 0) return cudaErrorInvalidValue;
        g.n_vtiles = ceil_div(g.nvfo, kNV);
        maxA = std::max(maxA, g.A);
        total += (double)g.n_vtiles * g.n_ttiles * g.A;
    }
    // Experiment knob (SDRPP_S1T_TILES_PER_CTA, default 1 = off): cap the grid so that a CTA gets at least that many time
    // tiles, leaving SMs to the tail / spectrum / ingest kernels of the neighbouring blocks. Measured on 16 .. 128 VFOs
    // (profiles/README.md): the step does not move for 1 .. 8 tiles per CTA -- a small VFO set is bound by the latency of
    // the tail chain, not by SMs -- and gets slower beyond.
    {
        double tiles = 0.0;
        int vtiles = 0;
        for (int i = 0; i < a.ngroups; i++) { tiles += (double)a.g[i].n_vtiles * a.g[i].n_ttiles; vtiles += a.g[i].n_vtiles; }
        static const int per_cta = getenv("SDRPP_S1T_TILES_PER_CTA") ? std::max(1, atoi(getenv("SDRPP_S1T_TILES_PER_CTA"))) : 1;
        const int want = std::max(vtiles, (int)std::ceil(tiles / per_cta));
        if (want < num_sms) num_sms = want;
    }
    // CTAs per VFO tile: proportional share of the SMs, at least 1, at most one per time tile
    int used = 0;
    for (int i = 0; i < a.ngroups; i++) {
        S1TGroupArgs& g = a.g[i];
        const double share = (double)num_sms * ((double)g.n_vtiles * g.n_ttiles * g.A) / total;
        g.cta_per_vtile = std::max(1, std::min(g.n_ttiles, (int)(share / g.n_vtiles)));
        used += g.cta_per_vtile * g.n_vtiles;
    }
    for (;;) { // hand the remaining SMs to the group with the longest per-CTA run
        int best = -1;
        double best_load = 0.0;
        for (int i = 0; i < a.ngroups; i++) {
            const S1TGroupArgs& g = a.g[i];
            if (g.cta_per_vtile >= g.n_ttiles || used + g.n_vtiles > num_sms) continue;
            const double load = (double)ceil_div(g.n_ttiles, g.cta_per_vtile) * g.A;
            if (load > best_load) { best_load = load; best = i; }
        }
        if (best < 0) break;
        a.g[best].cta_per_vtile++;
        used += a.g[best].n_vtiles;
    }
    int ctas = 0;
    for (int i = 0; i < a.ngroups; i++) { a.g[i].cta_begin = ctas; ctas += a.g[i].cta_per_vtile * a.g[i].n_vtiles; }
    // A-operand ring: as many 16 KB chunks as fit beside the largest B image
    // SDRPP_S1T_SMEM_CAP (experiment knob): a smaller budget shortens the operand ring; four chunks (190000) measured as fast
    // as the six to eight that fit in 227 KB
    static const size_t cap = getenv("SDRPP_S1T_SMEM_CAP") ? (size_t)atol(getenv("SDRPP_S1T_SMEM_CAP")) : 232448;
    int nch = kMaxChunks;
    while (nch > 2 * NKH && s1t_smem_bytes(NKH, maxA, nch) > cap) nch--;
    if (s1t_smem_bytes(NKH, maxA, nch) > 232448) return cudaErrorInvalidValue;   // `cap` is a preference, 227 KB the limit (A = 8)
    a.nchunks = nch;
    const size_t smem = s1t_smem_bytes(NKH, maxA, nch);
    if (cudaError_t e = ensure_dynamic_smem(NKH == 1 ? (const void*)s1t_kernel<1> : (const void*)s1t_kernel<2>, smem); e != cudaSuccess) return e;
    const S1TArgs* d = L.push(a);
    if (!d) return cudaErrorMemoryAllocation;
    return L.kernel(sid, NKH == 1 ? (const void*)s1t_kernel<1> : (const void*)s1t_kernel<2>, dim3((unsigned)ctas), dim3(kThreads), smem, d);
}

} // namespace sdrpp

#ifdef SDRPP_S1T_TRACE
extern "C" __attribute__((visibility("default"))) int sdrpp_cuda_debug_s1t_trace(long long* out, int rows) {
    if (rows > 256) rows = 256;
    return (int)cudaMemcpyFromSymbol(out, sdrpp::g_s1t_trace, sizeof(long long) * 24 * (size_t)rows);
}
#endif
