// Batched multi-VFO channelizer. Replaces, for every VFO at once, the per-VFO thread of the
// reference: Splitter memcpy (dsp/routing/splitter.h:46-60) -> dsp::channel::RxVFO::process
// (dsp/channel/rx_vfo.h:89-100) = FrequencyXlator (frequency_xlator.h:43-50) -> RationalResampler
// (PowerDecimator cascade power_decimator.h:51-67 + PolyphaseResampler polyphase_resampler.h:69-99)
// -> channel FIR (filter/fir.h:62-83) -> demod front end (demod/quadrature.h:41-56, am.h:122,
// ssb.h:90-101).
//
// Stage 1 (stage1_kernel) carries ~95 % of the arithmetic at high decimation: the NCO and the
// first decimating FIR, which both run at the full input rate. Every VFO of a group reads the SAME
// untranslated samples from one shared-memory tile (one bulk/TMA copy per CTA); the per-VFO work is
// one complex multiply per sample by a short phasor table, the taps stay real and shared (constant
// memory, uniform operands), and the inner loop is register-blocked packed FP32 FMA (FFMA2).
//
// The remaining stages run at 1/D of the rate (tail_kernel, one CTA per VFO).
#include "common.cuh"
#include "design.h"
#include "kernels.h"
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <type_traits>
#include <vector>

#ifndef SDRPP_S1_R
#define SDRPP_S1_R 6
#define SDRPP_S1_W 8
#endif

namespace sdrpp {

// ---------------------------------------------------------------------------------------------
// mbarrier / bulk-copy PTX wrappers (sm_90+/sm_100a)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait suspends for a bounded time per call; a copy that never lands traps instead of hanging the GPU
    for (uint32_t spins = 0; !mbar_try_wait(bar, parity); spins++)
        if (spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// 8-byte asynchronous global -> shared copy (LDGSTS): many in flight per thread without holding registers
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float2 phasor_u64(uint64_t phase) {
    // phase in turns * 2^64 -> (cos, sin). The top 32 bits as a signed fraction of a half turn.
    const float x = (float)(int32_t)(phase >> 32) * 4.656612873077393e-10f; // * 2^-31
    float s, c;
    sincospif(x, &s, &c);
    return make_float2(c, s);
}

// ---------------------------------------------------------------------------------------------
// Stage 1: NCO + first decimating FIR for 32 VFOs x 8*R input rows per CTA.
//
// With k = a*D + p (a < A = ceil(T/D), p < D) and row q = m + a holding samples x[q*D + p]:
//     y[m] = sum_a E[m+a] * V[m+a][a],   V[q][a] = sum_p h[a*D+p] * (x[q*D+p] * F[p]),
//     F[p] = e^{j w p} (per VFO, D entries),  E[q] = e^{j phi(first sample of row q)}.
// The per-VFO complex work is ONE complex multiply per sample (x*F); the taps stay real and
// shared by every VFO of the plan, so they sit in constant memory and reach the FMA pipe as a
// uniform operand. Per input sample and VFO that is 4 + 2*T/D FMAs instead of the 4*T/D of a
// complex-tap filter (or 8 + 2*T/D of rotate-then-filter).
// A thread (lane = VFO) owns R consecutive rows and their R*A complex accumulators V; rows are
// the same for all lanes, so sample loads are shared-memory broadcasts. The A partial sums of an
// output live in up to two neighbouring warps and are combined through shared memory; a CTA
// therefore computes 8*R rows for 8*R-(A-1) outputs.
// ---------------------------------------------------------------------------------------------
constexpr int kS1PoolFloats = 8192;
__constant__ __align__(16) float c_s1_taps[kS1PoolFloats];

// Pool entry: the plan's first FIR stored twice, zero-padded to A*D: at `off` as is, and at `off + A*D` with one
// leading zero tap. The second form lets the filter window start one sample earlier, which keeps the sample
// tile 16-byte aligned in the ring (a bulk-copy requirement) whatever the parity of the window start.
struct S1PoolEntry { int ratio, off, T, D, A; };
static std::vector<S1PoolEntry> g_s1_pool;
static std::mutex g_s1_mtx;
static bool g_s1_uploaded[64] = { false };

bool stage1_supported(int A, int D) { return D >= 2 && (D & (D - 1)) == 0 && D <= 128 && A >= 2 && A <= 7; }
int stage1_A(int T, int D) { return ceil_div(T + 1, D); } // room for the optional leading zero tap

// Builds the tap pool (every PowerDecimator plan's first FIR) and uploads it to the current device's constant
// memory. Returns the pool offset of the un-shifted taps for `ratio` (shifted form at +A*D), or -1.
int stage1_tap_offset(int ratio) {
    std::lock_guard<std::mutex> lck(g_s1_mtx);
    if (g_s1_pool.empty()) {
        int off = 0;
        for (int k = 1; k <= 13; k++) {
            std::vector<DecimStage> st = decim_plan(1 << k);
            if (st.empty()) continue;
            const int T = st[0].ntaps, D = st[0].decimation, A = stage1_A(T, D);
            if (!stage1_supported(A, D)) continue;
            g_s1_pool.push_back({ 1 << k, off, T, D, A });
            off += 2 * A * D;
        }
    }
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
    if (!g_s1_uploaded[dev]) {
        std::vector<float> pool(kS1PoolFloats, 0.0f);
        for (const S1PoolEntry& e : g_s1_pool) {
            if (e.off + 2 * e.A * e.D > kS1PoolFloats) return -1;
            std::vector<DecimStage> st = decim_plan(e.ratio);
            // stored pair-major: [pp][aa] = (h[aa*D+2pp], h[aa*D+2pp+1]), so the A tap pairs one sample pair needs
            // are contiguous in the constant bank (one address computation, immediate offsets per tap)
            auto put = [&](int base, int k, float v) {
                const int aa = k / e.D, p = k % e.D;
                pool[(size_t)(base + ((p >> 1) * e.A + aa) * 2 + (p & 1))] = v;
            };
            for (int k = 0; k < e.T; k++) {
                put(e.off, k, st[0].taps[k]);
                put(e.off + e.A * e.D, k + 1, st[0].taps[k]);
            }
        }
        if (cudaMemcpyToSymbol(c_s1_taps, pool.data(), sizeof(float) * kS1PoolFloats) != cudaSuccess) return -1;
        if (cudaStreamSynchronize(cudaStreamLegacy) != cudaSuccess) return -1; // complete before any non-blocking stream reads it
        g_s1_uploaded[dev] = true;
    }
    for (const S1PoolEntry& e : g_s1_pool) if (e.ratio == ratio) return e.off;
    return -1;
}

size_t stage1_g_elems(int A, int D, int nvfo) {
    (void)A;
    return (size_t)ceil_div(nvfo, 32) * (size_t)(D / 2) * 32;
}
void stage1_g_index(int A, int D, int v, int p, size_t* idx4, int* half) {
    (void)A;
    *idx4 = ((size_t)(v / 32) * (D / 2) + (size_t)(p / 2)) * 32 + (size_t)(v % 32);
    *half = p & 1;
}

// ---------------------------------------------------------------------------------------------
// The kernel runs on FFMA2 (fma.rn.f32x2, sm_100): one instruction updates the (re, im) pair of an
// accumulator. FFMA2 takes a scalar (register or uniform register) broadcast to both halves as an
// operand, so
//     (w.re, w.im) = x.re * (F.re, F.im) + x.im * (-F.im, F.re),    acc += h * (w.re, w.im)
// need no operand shuffling: x.re / x.im are the halves of the broadcast sample load and h is the
// uniform tap (FFMA2 R, R.F32x2, UR.F32, R.F32x2 in SASS). Measured on B200 (tools/_mb microbenchmarks,
// profiles/README.md): FFMA2 does not raise the FMA peak (128 FMA/clk/SM either way) and every non-FMA
// instruction in an FFMA2 stream costs about one FFMA2 slot, so the inner loop is kept to the sample,
// phasor and tap loads: the decimation is a template parameter for the plan shapes that matter (row
// strides become immediates) and the partially filled last tap slab has its own loop (no branches).
// ---------------------------------------------------------------------------------------------
#ifdef SDRPP_S1_TRACE
// Debug build only (-DSDRPP_S1_TRACE): per-CTA phase timestamps (globaltimer ns) of the packed kernel.
constexpr int kS1TraceCap = 8192;
__device__ long long g_s1_trace[kS1TraceCap][6];
__device__ __forceinline__ long long gtime() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define S1_TRACE(slot)                                                                            \
    do {                                                                                          \
        if (threadIdx.x == 0) {                                                                   \
            const unsigned lin_ = blockIdx.y * gridDim.x + blockIdx.x;                            \
            if (lin_ < (unsigned)kS1TraceCap) g_s1_trace[lin_][slot] = gtime();                   \
        }                                                                                         \
    } while (0)
#else
#define S1_TRACE(slot) do {} while (0)
#endif

// DT: decimation as a compile-time constant (row strides become immediates), or 0 to take it from the arguments.
template <int A, int R, int W, int DT>
__global__ void __launch_bounds__(W * 32, 2)
stage1_kernel(const Stage1Args* __restrict__ ap) {
    const Stage1Args a = *ap;   // per-block arguments (launcher descriptor)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int ROWS = W * R;
    constexpr int OUT = ROWS - (A - 1);
    constexpr int NP = R + A - 1;
    constexpr int NT = W * 32;
    const int D = DT > 0 ? DT : a.D, DP = D >> 1;

    S1_TRACE(0);
#ifdef SDRPP_S1_TRACE
    if (threadIdx.x == 0) {
        unsigned sm_; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_));
        const unsigned lin_ = blockIdx.y * gridDim.x + blockIdx.x;
        if (lin_ < (unsigned)kS1TraceCap) g_s1_trace[lin_][5] = sm_;
    }
#endif
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
    float2* parts = reinterpret_cast<float2*>(smem_raw + 128);    // [W][NP][32] partial outputs (epilogue)
    float4* Fs = reinterpret_cast<float4*>(parts + W * NP * 32);  // [DP][32]: (F[2pp], F[2pp+1]) per lane
    float2* xs = reinterpret_cast<float2*>(Fs + DP * 32);         // [ROWS][D] samples, natural order

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * OUT;
    const int vb = blockIdx.y;

    const uint32_t f_bytes = (uint32_t)DP * 32u * 16u;
    const uint32_t r0 = a.ring_first + (uint32_t)m0 * (uint32_t)D;
    const int64_t abs0 = a.abs_first + (int64_t)m0 * D;
    const uint32_t total = (uint32_t)(ROWS * D);
    const uint32_t i0 = r0 & a.ring.mask;
    const bool bulk_x = ((i0 & 1u) == 0) && (i0 + total - 1u <= a.ring.mask) && (abs0 >= a.abs_valid);
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
        mbar_expect_tx(bar, f_bytes + (bulk_x ? total * 8u : 0u));
        bulk_g2s(Fs, a.G + (size_t)vb * DP * 32, f_bytes, bar);
        if (bulk_x) bulk_g2s(xs, a.ring.base + i0, total * 8u, bar);
    }
    if (!bulk_x) {
        for (uint32_t n = tid; n < total; n += NT) {
            float2 s = a.ring.base[(r0 + n) & a.ring.mask];
            if (abs0 + (int64_t)n < a.abs_valid) s = make_float2(0.0f, 0.0f);
            xs[n] = s;
        }
    }
    __syncthreads(); // element-loaded tile complete; barrier init visible
    mbar_wait(bar, 0);
    S1_TRACE(1);

    float2 acc[R][A];
#pragma unroll
    for (int r = 0; r < R; r++)
#pragma unroll
        for (int aa = 0; aa < A; aa++) acc[r][aa] = make_float2(0.0f, 0.0f);

    const float4* xrow = reinterpret_cast<const float4*>(xs + (size_t)warp * R * D);
    const int rem = a.T - (A - 1) * D;  // taps in the last slab of the tap matrix
    const int pl = (rem + 1) >> 1;      // sample pairs that reach into the last slab; the others use A-1 slabs
    const float2* taps2 = reinterpret_cast<const float2*>(c_s1_taps + a.tap_off); // pair-major [pp][aa] (h[aa*D+2pp], h[aa*D+2pp+1])

    // one sample pair of every row of the warp against NA slabs of the tap matrix
    auto pair_step = [&](int pp, auto na_tag) {
        constexpr int NA = decltype(na_tag)::value;
        const float4 f = Fs[pp * 32 + lane];
        const float2 f0 = make_float2(f.x, f.y), g0 = make_float2(-f.y, f.x); // F[2pp], j*F[2pp]
        const float2 f1 = make_float2(f.z, f.w), g1 = make_float2(-f.w, f.z);
        float2 w0[R], w1[R];
#pragma unroll
        for (int r = 0; r < R; r++) {
            const float4 x = xrow[r * DP + pp]; // warp-uniform address: broadcast
            w0[r] = __ffma2_rn(make_float2(x.x, x.x), f0, __fmul2_rn(make_float2(x.y, x.y), g0));
            w1[r] = __ffma2_rn(make_float2(x.z, x.z), f1, __fmul2_rn(make_float2(x.w, x.w), g1));
        }
        const float2* tp = taps2 + pp * A;
#pragma unroll
        for (int aa = 0; aa < NA; aa++) {
            const float2 h = tp[aa]; // uniform constant load; FFMA2 broadcasts the scalar to both halves
#pragma unroll
            for (int r = 0; r < R; r++) {
                acc[r][aa] = __ffma2_rn(make_float2(h.x, h.x), w0[r], acc[r][aa]);
                acc[r][aa] = __ffma2_rn(make_float2(h.y, h.y), w1[r], acc[r][aa]);
            }
        }
    };
    {
        int pp = 0;
#pragma unroll 1
        for (; pp < pl; pp++) pair_step(pp, std::integral_constant<int, A>());
#pragma unroll 1
        for (; pp < DP; pp++) pair_step(pp, std::integral_constant<int, A - 1>());
    }

    S1_TRACE(2);
    const int v = vb * 32 + lane;
    const bool live = v < a.nvfo;
    VfoDev vd{};
    if (live) vd = a.vfos[v];
    // Rotate each row's sums by the NCO phase E[q] of the row's first sample and add them into the partial outputs
    // (output j = r - aa of this warp's rows, stored at j + A - 1). E comes from the 64-bit phase accumulator for the
    // warp's first row and advances by one row (D samples) per step: acc*E = acc.re*(E.re, E.im) + acc.im*(-E.im, E.re).
    float2 part[NP];
#pragma unroll
    for (int j = 0; j < NP; j++) part[j] = make_float2(0.0f, 0.0f);
    {
        const int q0 = m0 + warp * R;
        const uint64_t ph0 = vd.phi_ref + (uint64_t)(a.abs_first + (int64_t)q0 * D - vd.n_ref) * vd.dphi;
        float2 e = phasor_u64(ph0);
        const float2 step = phasor_u64((uint64_t)D * vd.dphi);
#pragma unroll
        for (int r = 0; r < R; r++) {
            if (r > 0) e = cmul(e, step);
            const float2 ej = make_float2(-e.y, e.x);
#pragma unroll
            for (int aa = 0; aa < A; aa++) {
                float2& pj = part[r - aa + (A - 1)];
                pj = __ffma2_rn(make_float2(acc[r][aa].x, acc[r][aa].x), e, pj);
                pj = __ffma2_rn(make_float2(acc[r][aa].y, acc[r][aa].y), ej, pj);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < NP; j++) parts[(warp * NP + j) * 32 + lane] = part[j];
    __syncthreads();
    S1_TRACE(3);

    if (live) {
        float2* __restrict__ out = vd.slab + a.out_off;
        for (int o = warp; o < OUT; o += W) {
            const int m = m0 + o;
            if (m >= a.M) break;
            // rows o .. o+A-1 belong to warps o/R .. (o+A-1)/R
            float2 y = make_float2(0.0f, 0.0f);
            const int w1 = o / R, w2 = (o + A - 1) / R;
            for (int w = w1; w <= w2 && w < W; w++) {
                const int j = o - w * R + (A - 1);
                if (j >= 0 && j < NP) {
                    const float2 t = parts[(w * NP + j) * 32 + lane];
                    y.x += t.x; y.y += t.y;
                }
            }
            out[m] = y;
        }
    }
    S1_TRACE(4);
}

template <int A, int R, int W, int DT>
static cudaError_t launch_stage1_t(Launcher& L, int sid, const Stage1Args& a) {
    constexpr int ROWS = W * R, OUT = ROWS - (A - 1), NP = R + A - 1;
    const size_t tile = (size_t)ROWS * a.D * sizeof(float2);
    const size_t parts = (size_t)W * NP * 32 * sizeof(float2);
    const size_t smem = 128 + parts + (size_t)(a.D / 2) * 32 * 16 + tile;
    if (cudaError_t e = ensure_dynamic_smem((const void*)stage1_kernel<A, R, W, DT>, smem); e != cudaSuccess) return e;
    dim3 grid(ceil_div(a.M, OUT), ceil_div(a.nvfo, 32));
    const Stage1Args* d = L.push(a);
    if (!d) return cudaErrorMemoryAllocation;
    return L.kernel(sid, (const void*)stage1_kernel<A, R, W, DT>, grid, dim3(W * 32), smem, d);
}

cudaError_t launch_stage1(Launcher& L, int sid, const Stage1Args& a) {
    if (a.M <= 0 || a.nvfo <= 0) return cudaSuccess;
    if (a.tap_off < 0) return cudaErrorInvalidValue;
    // the shapes the PowerDecimator plans produce at high ratios get the decimation as a compile-time constant
#define SDRPP_S1P_CASE(AA, DD) if (a.A == AA && a.D == DD) return launch_stage1_t<AA, SDRPP_S1_R, SDRPP_S1_W, DD>(L, sid, a);
    SDRPP_S1P_CASE(5, 32)
    SDRPP_S1P_CASE(5, 64)
    SDRPP_S1P_CASE(6, 64)
    SDRPP_S1P_CASE(7, 64)
    SDRPP_S1P_CASE(6, 128)
#undef SDRPP_S1P_CASE
    switch (a.A) {
    case 2: return launch_stage1_t<2, SDRPP_S1_R, SDRPP_S1_W, 0>(L, sid, a);
    case 3: return launch_stage1_t<3, SDRPP_S1_R, SDRPP_S1_W, 0>(L, sid, a);
    case 4: return launch_stage1_t<4, SDRPP_S1_R, SDRPP_S1_W, 0>(L, sid, a);
    case 5: return launch_stage1_t<5, SDRPP_S1_R, SDRPP_S1_W, 0>(L, sid, a);
    case 6: return launch_stage1_t<6, SDRPP_S1_R, SDRPP_S1_W, 0>(L, sid, a);
    case 7: return launch_stage1_t<7, SDRPP_S1_R, SDRPP_S1_W, 0>(L, sid, a);
    }
    return cudaErrorInvalidValue;
}

// D = 1: translation only (RationalResampler modes RESAMP_ONLY / NONE, rational_resampler.h:83-97)
__global__ void __launch_bounds__(256)
mix_only_kernel(const Stage1Args* __restrict__ ap) {
    const Stage1Args a = *ap;   // per-block arguments (launcher descriptor)
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (m >= a.M) return;
    const VfoDev vd = a.vfos[v];
    float2 x = a.ring.base[(a.ring_first + (uint32_t)m) & a.ring.mask];
    if (a.abs_first + m < a.abs_valid) x = make_float2(0.0f, 0.0f);
    const uint64_t ph = vd.phi_ref + (uint64_t)(a.abs_first + (int64_t)m - vd.n_ref) * vd.dphi;
    vd.slab[a.out_off + m] = cmul(x, phasor_u64(ph));
}

cudaError_t launch_mix_only(Launcher& L, int sid, const Stage1Args& a) {
    if (a.M <= 0 || a.nvfo <= 0) return cudaSuccess;
    dim3 grid(ceil_div(a.M, 256), a.nvfo);
    const Stage1Args* d = L.push(a);
    if (!d) return cudaErrorMemoryAllocation;
    return L.kernel(sid, (const void*)mix_only_kernel, grid, dim3(256), 0, d);
}

// ---------------------------------------------------------------------------------------------
// Tail: remaining decimating FIRs, polyphase resampler, channel FIR, demod front end.
// One CTA per VFO; stages run back to back on the VFO's slab (L2-resident), __syncthreads between.
// ---------------------------------------------------------------------------------------------
constexpr int kTailThreads = 256;
// 36 KB samples + 4 KB slack + 12 KB taps = 52 KB: four CTAs per SM, so 512 VFOs are a single wave on 148 SMs
#ifndef SDRPP_TAIL_SMEM_SAMPLES
#define SDRPP_TAIL_SMEM_SAMPLES 4608
#endif
constexpr int kTailSmemSamples = SDRPP_TAIL_SMEM_SAMPLES; // staged input samples per chunk
constexpr int kTailTapFloats = 3072;   // staged taps per stage

// Stage one chunk of a stage's input [first, first+n) from the slab (L2) into shared memory.
// Decimating FIRs store it transposed by D -- element i at [i % D][i / D] with an odd row stride -- so
// that consecutive outputs (lanes) read consecutive addresses for every tap; others keep natural order.
__device__ __forceinline__ void tail_stage_in(float2* sm, const float2* src, int n, int D, int qs, bool wait = true) {
    const int lg = 31 - __clz(D); // D is a power of two (every PowerDecimator stage decimates by 2, 4, 8, ...)
    for (int i = threadIdx.x; i < n; i += kTailThreads) cp_async8(sm + (i & (D - 1)) * qs + (i >> lg), src + i);
    if (wait) cp_async_wait_all();
}

// FIR / decimating FIR stage, register-blocked: a thread owns OB = 4 consecutive outputs, so a staged sample is
// loaded once for up to four taps (the stage was shared-memory-bandwidth-bound at one load per tap and output,
// profiles/r1d). With M = OB*D, input element i of the chunk sits at plane i % M, position i / M; output 4t+r, tap k
// reads element M*t + s with s = r*D + k, so for every s the threads of a warp read consecutive positions of one
// plane, and the four taps that meet that sample, (h[s], h[s-D], h[s-2D], h[s-3D]), come from one broadcast load of
// a table built per stage. Few outputs and many taps (the channel filter): the s range is split over KS thread groups
// and reduced through shared memory. out[o] = sum_k buf[offset + o*D + k] * h[k] (decimating_fir.h:45-68, fir.h:62-83).
template <int NT>
__device__ __forceinline__ void tail_fir_blocked(const TailStage& st, int D, const float2* __restrict__ buf, float2* __restrict__ out,
                                                 float2* tsm, float4* tt) {
    constexpr int OB = 4;
    constexpr int kTabEntries = kTailTapFloats / 4;
    const int tid = threadIdx.x;
    const int T = st.T, M = OB * D, lgM = 31 - __clz(M);
    const int S = (OB - 1) * D + T, SQ = (S + M - 1) >> lgM;
    int ch = min(OB * NT, (kTailSmemSamples - T - 16 * D) / D) & ~(OB - 1);
    if (ch < OB) ch = OB;
    for (int o0 = 0; o0 < st.n_out; o0 += ch) {
        const int co = min(ch, st.n_out - o0);
        const int nthr = (co + OB - 1) / OB;
        const int qs = (nthr + SQ) | 1;
        const int n = (co - 1) * D + T; // staged samples that exist
        const float2* __restrict__ src = buf + st.offset + o0 * D;
        {
            // asynchronous element copies (the copy is L2-latency-bound: everything a thread moves is in flight at once)
            const int total = M * qs;
            for (int i = tid; i < total; i += NT) {
                float2* dst = tsm + (i & (M - 1)) * qs + (i >> lgM);
                if (i < n) cp_async8(dst, src + i);
                else *dst = make_float2(0.0f, 0.0f);
            }
            // not waited for here: the first tap table below is built while the copies are in flight
        }
        int KS = 1;
        while (KS < 8 && nthr * KS * 2 <= NT && SQ >= 8 * KS * 2) KS *= 2;
        const int Sk = (SQ + KS - 1) / KS;                  // rows of M s-values per thread group
        const int SEGq = max(1, min(Sk, kTabEntries / (KS * M)));
        const int t_out = tid % nthr, ks = tid / nthr;
        const bool active = ks < KS;
        float2 acc[OB];
#pragma unroll
        for (int r = 0; r < OB; r++) acc[r] = make_float2(0.0f, 0.0f);
        for (int q0 = 0; q0 < Sk; q0 += SEGq) {
            // tap table of this step: for each group, rows [ks*Sk + q0, +SEGq)
            for (int e = tid; e < KS * SEGq * M; e += NT) {
                const int g = e / (SEGq * M), loc = e - g * (SEGq * M);
                const int row = g * Sk + q0 + (loc >> lgM);
                float4 h = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                if (row < min((g + 1) * Sk, SQ) && (loc >> lgM) < Sk - q0) {
                    const int sv = (row << lgM) + (loc & (M - 1));
                    h.x = (sv < T) ? __ldg(st.taps + sv) : 0.0f;
                    h.y = (sv - D >= 0 && sv - D < T) ? __ldg(st.taps + sv - D) : 0.0f;
                    h.z = (sv - 2 * D >= 0 && sv - 2 * D < T) ? __ldg(st.taps + sv - 2 * D) : 0.0f;
                    h.w = (sv - 3 * D >= 0 && sv - 3 * D < T) ? __ldg(st.taps + sv - 3 * D) : 0.0f;
                }
                tt[e] = h;
            }
            if (q0 == 0) cp_async_wait_all();
            __syncthreads(); // samples (first step) and the table are staged
            if (active) {
                const int r0 = ks * Sk + q0;
                const int r1 = min(min(r0 + SEGq, (ks + 1) * Sk), SQ);
                for (int row = r0; row < r1; row++) {
                    const float2* __restrict__ xp = tsm + t_out + row;
                    const float4* __restrict__ tp = tt + ((ks * SEGq + (row - r0)) << lgM);
                    // four window steps per trip, all eight loads first: left to the compiler (#pragma unroll 4) every step's
                    // FMAs waited for that step's own two loads in the same registers (SASS), a latency chain per step
                    for (int sp = 0; sp < M; sp += 4) {     // M = OB * D is a multiple of 4
                        const float2 v0 = xp[sp * qs], v1 = xp[(sp + 1) * qs], v2 = xp[(sp + 2) * qs], v3 = xp[(sp + 3) * qs];
                        const float4 h0 = tp[sp], h1 = tp[sp + 1], h2 = tp[sp + 2], h3 = tp[sp + 3];
                        // packed FMA: (re, im) of one accumulator per instruction, the tap broadcast to both halves
                        acc[0] = __ffma2_rn(make_float2(h0.x, h0.x), v0, acc[0]);
                        acc[1] = __ffma2_rn(make_float2(h0.y, h0.y), v0, acc[1]);
                        acc[2] = __ffma2_rn(make_float2(h0.z, h0.z), v0, acc[2]);
                        acc[3] = __ffma2_rn(make_float2(h0.w, h0.w), v0, acc[3]);
                        acc[0] = __ffma2_rn(make_float2(h1.x, h1.x), v1, acc[0]);
                        acc[1] = __ffma2_rn(make_float2(h1.y, h1.y), v1, acc[1]);
                        acc[2] = __ffma2_rn(make_float2(h1.z, h1.z), v1, acc[2]);
                        acc[3] = __ffma2_rn(make_float2(h1.w, h1.w), v1, acc[3]);
                        acc[0] = __ffma2_rn(make_float2(h2.x, h2.x), v2, acc[0]);
                        acc[1] = __ffma2_rn(make_float2(h2.y, h2.y), v2, acc[1]);
                        acc[2] = __ffma2_rn(make_float2(h2.z, h2.z), v2, acc[2]);
                        acc[3] = __ffma2_rn(make_float2(h2.w, h2.w), v2, acc[3]);
                        acc[0] = __ffma2_rn(make_float2(h3.x, h3.x), v3, acc[0]);
                        acc[1] = __ffma2_rn(make_float2(h3.y, h3.y), v3, acc[1]);
                        acc[2] = __ffma2_rn(make_float2(h3.z, h3.z), v3, acc[2]);
                        acc[3] = __ffma2_rn(make_float2(h3.w, h3.w), v3, acc[3]);
                    }
                }
            }
            __syncthreads(); // table (and after the last step the samples) may be overwritten
        }
        if (KS > 1) {
            float2* red = tsm; // [ks][t_out][OB]
            if (active) {
#pragma unroll
                for (int r = 0; r < OB; r++) red[(ks * nthr + t_out) * OB + r] = acc[r];
            }
            __syncthreads();
            if (ks == 0) {
                for (int g = 1; g < KS; g++) {
#pragma unroll
                    for (int r = 0; r < OB; r++) { const float2 v = red[(g * nthr + t_out) * OB + r]; acc[r].x += v.x; acc[r].y += v.y; }
                }
            }
        }
        if (ks == 0) {
#pragma unroll
            for (int r = 0; r < OB; r++) {
                const int o = OB * t_out + r;
                if (o < co) out[o0 + o] = acc[r];
            }
        }
        __syncthreads();
    }
}

// Radio IF chain on one block of a VFO's output, in place on p[0..n) (p[-1] receives the previous block's last
// output). Both blocks are sequential in the reference and cheap at the VFO's output rate, so one thread runs them
// with the reference's operation order (_rn intrinsics: no contraction):
//   dsp::noise_reduction::NoiseBlanker::process (noise_reduction/noise_blanker.h:39-59)
//   dsp::noise_reduction::Squelch::process (noise_reduction/squelch.h:34-64); its block counter is a function-local
//   static there (one counter for every instance in the process) and per VFO here.
__device__ __noinline__ void if_chain_run(float* __restrict__ r, float2* p, int n) {
    if (r[IF_NB_ON] != 0.0f) {
        const float rate = r[IF_NB_RATE], inv_rate = r[IF_NB_INVRATE], level = r[IF_NB_LEVEL];
        float amp = r[IF_NB_AMP];
        for (int i = 0; i < n; i++) {
            const float2 v = p[i];
            const float in_amp = __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y)));
            float gain = 1.0f;
            if (in_amp != 0.0f) {
                amp = __fadd_rn(__fmul_rn(amp, inv_rate), __fmul_rn(in_amp, rate));
                const float excess = __fdiv_rn(in_amp, amp);
                if (excess > level) gain = __fdiv_rn(1.0f, excess);
            }
            p[i] = make_float2(__fmul_rn(v.x, gain), __fmul_rn(v.y, gain));
        }
        r[IF_NB_AMP] = amp;
    }
    if (r[IF_SQ_ON] != 0.0f) {
        float sum = 0.0f; // volk_32fc_magnitude_32f + volk_32f_accumulator_s32f (generic: in order)
        for (int i = 0; i < n; i++) {
            const float2 v = p[i];
            sum = __fadd_rn(sum, __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y))));
        }
        sum = __fdiv_rn(sum, (float)n);
        const float level = __fmul_rn(20.0f, log10f(sum));
        const float thr = r[IF_SQ_LEVEL];
        bool mute = r[IF_SQ_MUTE] != 0.0f;
        int cnt = (int)r[IF_SQ_CNT];
        if (mute) {
            if (level < thr || cnt <= 0) cnt = 10;
            else if (--cnt == 0) mute = false;
        } else if (level < __fadd_rn(thr, -1.0f)) {
            cnt = 0;
            mute = true;
        }
        r[IF_SQ_MUTE] = mute ? 1.0f : 0.0f;
        r[IF_SQ_CNT] = (float)cnt;
        r[IF_SQ_LAST_DB] = level;
        if (mute) for (int i = 0; i < n; i++) p[i] = make_float2(0.0f, 0.0f);
    }
    p[-1] = make_float2(r[IF_PREV_RE], r[IF_PREV_IM]);
}

static_assert(SDRPP_TAIL_SMEM_SAMPLES != 4608 || kIfMaxBlock == kTailSmemSamples / 2 - 2 * kIfMaxBins, "IF chain staging: two half areas with room for the FMIF history");

// dsp::noise_reduction::FMIF::process (noise_reduction/fm_if.h:45-74) for a whole block, one output per thread: the
// last `bins` samples under a Nuttall window, forward DFT (bins = 9, 15, 31, 32 in the radio module: direct
// evaluation, every thread of a warp is at the same (bin, sample) pair so window and twiddle loads are broadcasts),
// first strongest bin, and element bins/2 of the backward DFT of that single bin = X[idx] * e^{+2 pi j idx (bins/2) / bins}.
// p[-(bins-1) .. n) holds history | block on return of the history load; q receives the outputs, q[-1] the previous one.
__device__ __noinline__ void if_fmif_run(float* __restrict__ r, int bins, float2* p, float2* q, int n, float* scratch) {
    const int tid = threadIdx.x;
    float* win = scratch;                                        // [bins]
    float2* tw = reinterpret_cast<float2*>(scratch + kIfMaxBins); // [bins]
    for (int i = tid; i < bins; i += kTailThreads) {
        win[i] = r[IF_WIN + i];
        tw[i] = make_float2(r[IF_TW + 2 * i], r[IF_TW + 2 * i + 1]);
    }
    for (int i = tid; i < bins - 1; i += kTailThreads) p[i - (bins - 1)] = make_float2(r[IF_HIST + 2 * i], r[IF_HIST + 2 * i + 1]);
    if (tid == 0) q[-1] = make_float2(r[IF_PREV_RE], r[IF_PREV_IM]); // the chain's previous output
    __syncthreads();
    const int half = bins / 2;
    for (int i = tid; i < n; i += kTailThreads) {
        const float2* x = p + i - (bins - 1);
        float best_amp = 0.0f;
        float2 best = make_float2(0.0f, 0.0f);
        int best_b = 0;
        for (int b = 0; b < bins; b++) {
            float re = 0.0f, im = 0.0f;
            int m = 0; // (b * k) mod bins
            for (int k = 0; k < bins; k++) {
                const float2 v = x[k];
                const float w = win[k];
                const float2 t = tw[m];
                const float xr = __fmul_rn(v.x, w), xi = __fmul_rn(v.y, w);
                re = fmaf(xr, t.x, fmaf(-xi, t.y, re));
                im = fmaf(xr, t.y, fmaf(xi, t.x, im));
                m += b; if (m >= bins) m -= bins;
            }
            const float amp = __fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
            if (b == 0 || amp > best_amp) { best_amp = amp; best = make_float2(re, im); best_b = b; }
        }
        const float2 t = tw[(best_b * half) % bins]; // backward transform: conjugate twiddle
        q[i] = make_float2(best.x * t.x + best.y * t.y, best.y * t.x - best.x * t.y);
    }
    __syncthreads();
    // the last bins-1 input samples are the next block's history (fm_if.h:72)
    for (int i = tid; i < bins - 1; i += kTailThreads) {
        const float2 v = p[n - (bins - 1) + i];
        r[IF_HIST + 2 * i] = v.x; r[IF_HIST + 2 * i + 1] = v.y;
    }
    __syncthreads();
}

// Demod front end of one output sample y (prev = the sample before it), in the reference's operation order with no
// contraction: Quadrature (demod/quadrature.h:41-56: (y * conj(prev)).phase() * invDeviation with complex_t::operator*,
// dsp/types.h:23-25), AM magnitude (volk_32fc_magnitude_32f, am.h:122: bit-exact with IEEE sqrt), SSB second
// translation + real part (ssb.h:90-101).
__device__ __forceinline__ float demod_front_end(const TailGroup& g, const VfoDev& vd, float2 y, float2 p, int i) {
    if (g.demod == 1) {
        const float dre = __fadd_rn(__fmul_rn(y.x, p.x), __fmul_rn(y.y, p.y));
        const float dim = __fsub_rn(__fmul_rn(y.y, p.x), __fmul_rn(y.x, p.y));
        return __fmul_rn(atan2f(dim, dre), g.inv_dev);
    }
    if (g.demod == 2) return __fsqrt_rn(__fadd_rn(__fmul_rn(y.x, y.x), __fmul_rn(y.y, y.y)));
    const float2 w = phasor_u64((uint64_t)(g.abs_out + i) * vd.dphi2);
    return __fsub_rn(__fmul_rn(y.x, w.x), __fmul_rn(y.y, w.y));
}

__global__ void __launch_bounds__(kTailThreads, 4)
tail_kernel(const TailArgs* __restrict__ ap) {
    const TailArgs& a = *ap;   // per-block arguments in the launcher's descriptor (device memory)
    extern __shared__ __align__(16) unsigned char tail_smem[];
    float* ttaps = reinterpret_cast<float*>(tail_smem);                           // [kTailTapFloats]
    float2* tsm = reinterpret_cast<float2*>(tail_smem + kTailTapFloats * 4);      // staged samples
    int vi = blockIdx.x, gi = 0;
    while (gi < a.ngroups - 1 && vi >= a.g[gi].nvfo) { vi -= a.g[gi].nvfo; gi++; }
    const int tid = threadIdx.x;
    // The group record lives in the per-block descriptor (device memory) since the command-list engine: read through a
    // reference into global memory, every st.T / st.taps / st.n_out was a global load again after each store (the compiler
    // cannot rule out aliasing with the slab). One coalesced copy into shared memory, like tail_fast_kernel.
    __shared__ TailGroup sg;
    {
        const int* src = reinterpret_cast<const int*>(&a.g[gi]);
        int* dst = reinterpret_cast<int*>(&sg);
        for (int i = tid; i < (int)(sizeof(TailGroup) / sizeof(int)); i += kTailThreads) dst[i] = src[i];
    }
    const VfoDev vd = a.vfos[a.g[gi].first_vfo + vi];
    __syncthreads();
    const TailGroup& g = sg;
    float2* slab = vd.slab;

    for (int s = g.s_begin; s < g.nstages; s++) { // stage 0 may already have run in tail_stage0_wide_kernel
        const TailStage& st = g.st[s];
        const int T = st.T, hist = T - 1;
        float2* buf = slab + st.in_off - hist; // [hist | n_in]
        float2* out = slab + ((s + 1 < g.nstages) ? g.st[s + 1].in_off : g.final_off);
        if (st.type != TAIL_POLY) {
            tail_fir_blocked<kTailThreads>(st, st.type == TAIL_DECFIR ? st.D : 1, buf, out, tsm, reinterpret_cast<float4*>(ttaps));
        } else {
            // polyphase resampler: consecutive outputs use different tap phases and input strides; one output per thread
            const bool taps_staged = (long long)st.interp * T <= kTailTapFloats;
            bool taps_pending = taps_staged; // staged while the first chunk's copies are in flight
            int ch = (int)(((long long)(kTailSmemSamples - T - 2) * st.interp) / st.D);
            if (ch >= kTailThreads) ch -= ch % kTailThreads;
            if (ch < 1) ch = 1;
            for (int o0 = 0; o0 < st.n_out; o0 += ch) {
                const int o1 = min(st.n_out, o0 + ch);
                const long long P0 = (long long)st.phase + (long long)o0 * st.D;
                const long long P1 = (long long)st.phase + (long long)(o1 - 1) * st.D;
                const int first = st.offset + (int)(P0 / st.interp);
                const int last = st.offset + (int)(P1 / st.interp) + T - 1;
                tail_stage_in(tsm, buf + first, last - first + 1, 1, 0, false);
                if (taps_pending) {
                    for (int i = tid; i < st.interp * T; i += kTailThreads) ttaps[i] = __ldg(st.taps + i);
                    taps_pending = false;
                }
                cp_async_wait_all();
                __syncthreads();
                for (int o = o0 + tid; o < o1; o += kTailThreads) {
                    // closed form of polyphase_resampler.h:75-93
                    const long long P = (long long)st.phase + (long long)o * st.D;
                    const float2* x = tsm + (st.offset + (int)(P / st.interp) - first);
                    const int ph = (int)(P % st.interp);
                    const float* h = taps_staged ? ttaps + ph * T : st.taps + (size_t)ph * T; // rows are not 16-byte aligned in general
                    float2 a0 = make_float2(0.0f, 0.0f), a1 = make_float2(0.0f, 0.0f);
                    int k = 0;
                    for (; k + 2 <= T; k += 2) {
                        const float2 v0 = x[k], v1 = x[k + 1];
                        const float t0 = h[k], t1 = h[k + 1];
                        a0 = __ffma2_rn(make_float2(t0, t0), v0, a0);
                        a1 = __ffma2_rn(make_float2(t1, t1), v1, a1);
                    }
                    if (k < T) { const float2 v = x[k]; const float t = h[k]; a0 = __ffma2_rn(make_float2(t, t), v, a0); }
                    out[o] = make_float2(a0.x + a1.x, a0.y + a1.y);
                }
                __syncthreads();
            }
        }
        // Stage 0 reads the double-buffered stage-1 region, so its history goes to the OTHER region (always, even
        // for an empty block); later stages shift in place.
        if (s == 0 || st.n_in > 0) {
            float2* dst = (s == 0) ? (slab + g.carry0_off - hist) : buf;
            float2 keep[8];
            int n = 0;
            for (int i = tid; i < hist && n < 8; i += kTailThreads, n++) keep[n] = buf[st.n_in + i];
            __syncthreads();
            n = 0;
            for (int i = tid; i < hist && n < 8; i += kTailThreads, n++) dst[i] = keep[n];
            __syncthreads();
        }
    }

    // final output, [radio IF chain], demod front end, results arena
    float2* fin = slab + g.final_off; // fin[-1] = last output of the previous block
    float2* o_iq = a.arena_iq + vd.out_off;
    float* o_dm = a.arena_demod + vd.out_off;
    // The demod front end reads `src`: the VFO output itself, or -- with an IF chain record -- a copy in shared
    // memory that the chain's blocks have worked on (the iq result stays the raw VFO output = vfo->output).
    const float2* src = fin;
    const bool if_on = vd.ifs != nullptr && g.n_final > 0 && g.n_final <= kIfMaxBlock; // the engine refuses larger blocks
    if (if_on) {
        float2* p = tsm + kIfMaxBins; // p[-1] = last IF-chain output of the previous block (or FMIF history in front)
        __syncthreads();
        for (int i = tid; i < g.n_final; i += kTailThreads) p[i] = fin[i];
        __syncthreads();
        if (tid == 0) if_chain_run(vd.ifs, p, g.n_final);
        __syncthreads();
        const int bins = (int)vd.ifs[IF_FMIF_BINS];
        if (bins > 0) {
            float2* q = p + kIfMaxBlock + kIfMaxBins; // second half of the staging area
            if_fmif_run(vd.ifs, bins, p, q, g.n_final, ttaps);
            p = q;
        }
        src = p;
    }
    for (int i = tid; i < g.n_final; i += kTailThreads) {
        o_iq[i] = fin[i];
        if (g.demod != 0) o_dm[i] = demod_front_end(g, vd, src[i], src[i - 1], i);
    }
    __syncthreads();
    if (tid == 0) {
        // previous-sample state of the quadrature demod; with no tail stage the final output IS the stage-1 region
        float2* nxt = (g.nstages == 0) ? (slab + g.carry0_off) : fin;
        if (g.n_final > 0) nxt[-1] = fin[g.n_final - 1];
        else if (g.nstages == 0) nxt[-1] = fin[-1];
        if (if_on) { vd.ifs[IF_PREV_RE] = src[g.n_final - 1].x; vd.ifs[IF_PREV_IM] = src[g.n_final - 1].y; }
    }
}

// ---------------------------------------------------------------------------------------------
// Tail, low-latency form. tail_kernel walks a VFO's stages through the slab: every stage stages its input from L2,
// builds a tap table, computes, stores, carries its history -- four to six dependent L2 round trips and as many
// barriers per stage, ~36 us per CTA whatever the VFO count. When a VFO's stage inputs of this block all fit in shared
// memory at once (every high-decimation plan: a few thousand samples), ONE round trip fetches everything -- the first
// stage's input, every stage's T-1 history samples, every tap table, the previous final sample -- the stages then run
// back to back out of shared memory with one barrier each, and only the results and the new histories go back.
// Same arithmetic (fir.h:62-83, decimating_fir.h:45-68, polyphase_resampler.h:69-99), same slab layout, so the two
// kernels can alternate block by block. The engine picks this one per group and block (tail_fast_fits).
// ---------------------------------------------------------------------------------------------
// Two shapes of the same kernel: 1024 threads and up to 160 KB of stage regions when there are fewer VFOs than SMs (one CTA
// per SM, all of its threads on one VFO), and 256 threads with <= 39.7 KB of regions -- four CTAs per SM, so that 512 VFOs
// are one wave and one CTA's round trip to L2 hides behind its neighbours' arithmetic -- when the VFO set is larger.
constexpr int kFastThreadsWide = 1024, kFastThreadsNarrow = 256;
__host__ __device__ constexpr int fast_max_samples(int nt) { return nt >= 1024 ? 20480 : 5080; }  // float2 of stage regions + split-K scratch
constexpr int kFastTapFloats = 4096;
constexpr int kFastOB = 4;              // outputs per thread of a FIR stage (register blocking)
// split-K partial sums: threads * kFastOB float2 of scratch

// A stage's input region [T-1 history | n_in data] in shared memory. FIR stages read it register-blocked (a thread owns
// four consecutive outputs, so a sample and a float4 of taps serve up to four MACs -- at one load per MAC the kernel was
// bound by shared-memory bandwidth, profiles/r2d), which needs the region transposed by M = 4*D: element i at plane
// i % M, position i / M, so that the threads of a warp (consecutive output groups) read consecutive positions of one
// plane. The polyphase stage and the final output keep natural order (M = 1).
struct FastRegion {
    int base;   // float2 offset of the region in the sample area
    int M, lgM; // transposition modulus (power of two) and its log2; M = 1: natural order
    int qs;     // positions per plane (odd), M > 1
    int len;    // elements (history + data)
};
__host__ __device__ inline int fast_stage_M(const TailStage& st) { return st.type == TAIL_POLY ? 1 : kFastOB * (st.type == TAIL_DECFIR ? st.D : 1); }
__host__ __device__ inline int fast_ilog2(int v) {
#ifdef __CUDA_ARCH__
    return v <= 1 ? 0 : 32 - __clz(v - 1);
#else
    int l = 0; while ((1 << l) < v) l++; return l;
#endif
}
__host__ __device__ inline int fast_region_floats2(const TailStage& st, FastRegion* r, int base) {
    const int len = (st.T - 1) + st.n_in;
    const int M = fast_stage_M(st);
    const int lgM = fast_ilog2(M);
    // the last output group of a blocked reader runs up to 6*D elements past the data (outputs beyond n_out, zero taps):
    // those elements exist and are zero (0 * garbage could be NaN)
    const int D = st.type == TAIL_DECFIR ? st.D : 1;
    const int padded = len + 6 * D + 8;
    const int qs = M > 1 ? (((padded + M - 1) >> lgM) + 1) | 1 : 0;
    if (r) { r->base = base; r->M = M; r->lgM = lgM; r->qs = qs; r->len = len; }
    return M > 1 ? ((M * qs + 1) & ~1) : ((padded + 1) & ~1);
}
// float4 table (h[s], h[s-D], h[s-2D], h[s-3D]) for s < T + 3D of a FIR stage, without / with the raw taps behind it
__host__ __device__ inline int fast_table_floats(const TailStage& st) {
    const int D = st.type == TAIL_DECFIR ? st.D : 1;
    return 4 * (st.T + (kFastOB - 1) * D);
}
__host__ __device__ inline int fast_tap_floats(const TailStage& st) {
    if (st.type == TAIL_POLY) return (st.interp * st.T + 3) & ~3;
    // the raw taps arrive by cp.async with everything else the block needs (one round trip) and the table is built from
    // them out of shared memory: built straight from global loads, five stages of load -> store chains cost 16 k cycles
    return fast_table_floats(st) + ((st.T + 3) & ~3);
}

__host__ __device__ inline int fast_tap_bytes(int tap_floats) { return (tap_floats * 4 + 15) & ~15; }
// The narrow shape has to fit beside the persistent stage-1 CTA of the NEXT block (189 KB of the SM's 228 KB with its
// four-chunk operand ring): tap tables + stage regions <= 41 KB, sized per group instead of a fixed tap area.
constexpr int kFastNarrowBytes = 41728;

bool tail_fast_fits(const TailGroup& g, int* samples, int threads, int* tap_floats) {
    const int kFastScratch = threads * kFastOB, kFastMaxSamples = fast_max_samples(threads);
    if (g.nstages > kTailMaxStages || g.nstages < 1) return false;
    int pos = 0, tp = 0;
    for (int s = g.s_begin; s < g.nstages; s++) {
        const TailStage& st = g.st[s];
        if (st.type == TAIL_DECFIR && (st.D & (st.D - 1)) != 0) return false;
        pos += fast_region_floats2(st, nullptr, 0);
        tp += fast_tap_floats(st);
    }
    pos += (g.n_final + 2 + 1) & ~1;
    pos += kFastScratch;
    if (samples) *samples = pos;
    if (tap_floats) *tap_floats = tp;
    if (threads < 1024 && fast_tap_bytes(tp) + (pos + 8) * (int)sizeof(float2) > kFastNarrowBytes) return false;
    return pos <= kFastMaxSamples && tp <= kFastTapFloats;
}

__device__ __forceinline__ float2& fast_at(float2* x, const FastRegion& r, int i) {
    return r.M > 1 ? x[r.base + (i & (r.M - 1)) * r.qs + (i >> r.lgM)] : x[r.base + i];
}

#ifdef SDRPP_TAILFAST_TRACE
__device__ long long g_tailfast_trace[16];
#define TF_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_tailfast_trace[i] = clock64(); } while (0)
#else
#define TF_MARK(i) do {} while (0)
#endif

template <int kFastThreads>
__global__ void __launch_bounds__(kFastThreads, kFastThreads >= 1024 ? 1 : 4)
tail_fast_kernel(const TailArgs* __restrict__ ap) {
    const TailArgs& a = *ap;   // per-block arguments in the launcher's descriptor (device memory)
    extern __shared__ __align__(16) unsigned char tail_smem[];
    float* taps = reinterpret_cast<float*>(tail_smem);                          // the group's tap tables, then (16-byte aligned):
    float2* x;                                                                  // stage regions ..., [prev | final], split-K scratch
    const int tid = threadIdx.x;
    int vi = blockIdx.x, gi = 0;
    while (gi < a.ngroups - 1 && vi >= a.g[gi].nvfo) { vi -= a.g[gi].nvfo; gi++; }
    TF_MARK(0);
    // The group record (480 bytes of a 2.9 KB kernel parameter) is read word by word by different threads, so its
    // constant-cache misses overlap, and lives in shared memory from here on: walked by one thread after the other it
    // cost 3.9 k cycles of cold misses before the first load was issued (tools/tailfast_trace.py).
    __shared__ TailGroup sg;
    {
        const int* src = reinterpret_cast<const int*>(&a.g[gi]);
        int* dst = reinterpret_cast<int*>(&sg);
        for (int i = tid; i < (int)(sizeof(TailGroup) / sizeof(int)); i += kFastThreads) dst[i] = src[i];
    }
    const VfoDev vd = a.vfos[a.g[gi].first_vfo + vi];
    __syncthreads();
    const TailGroup& g = sg;
    float2* slab = vd.slab;

    // Layout: thread s sizes stage s (the divisions and logarithms once, not 1024 times: 5.6 k cycles before), then everyone
    // takes the prefix sums.
    // Layout in SHARED memory. As per-thread arrays (dynamically indexed, so in local memory: 168 bytes x 1024 threads
    // against an L1 that the shared-memory carve-out leaves at 28 KB) every reg[s] / toff[s] access was a trip to L2 -- the
    // load-issue phase alone took 16 k cycles (tools/tailfast_trace.py). Thread s sizes stage s, thread 0 takes the prefix sums.
    __shared__ FastRegion reg[kTailMaxStages + 1];
    __shared__ int ssize[kTailMaxStages], toff[kTailMaxStages + 1];
    if (tid >= g.s_begin && tid < g.nstages) {
        ssize[tid] = fast_region_floats2(g.st[tid], &reg[tid], 0);
        toff[tid] = fast_tap_floats(g.st[tid]);
    }
    __syncthreads();
    if (tid == 0) {
        int pos0 = 0, tp = 0;
        for (int s = g.s_begin; s < g.nstages; s++) {
            const int t = toff[s];
            toff[s] = tp; tp += t;
            reg[s].base = pos0; pos0 += ssize[s];
        }
        toff[kTailMaxStages] = tp;
        FastRegion& f = reg[g.nstages];      // final: natural order, element 0 = last output of the previous block
        f.base = pos0; f.M = 1; f.lgM = 0; f.qs = 0; f.len = g.n_final + 1;
    }
    __syncthreads();
    x = reinterpret_cast<float2*>(tail_smem + fast_tap_bytes(toff[kTailMaxStages]));
    const FastRegion rf = reg[g.nstages];
    float2* fin = x + rf.base + 1;
    float2* scratch = x + rf.base + ((g.n_final + 2 + 1) & ~1);

    TF_MARK(1);
    // ---- one round trip: everything this block needs ------------------------------------------------------------------
    if (g.s_begin < g.nstages) {
        const TailStage& st = g.st[g.s_begin];
        const FastRegion r = reg[g.s_begin];   // by value: registers, not a shared-memory reload per access
        const float2* __restrict__ src = slab + st.in_off - (st.T - 1);     // [history | data], contiguous in the slab
        for (int i = tid; i < r.len; i += kFastThreads) cp_async8(&fast_at(x, r, i), src + i);
    } else {
        const float2* __restrict__ src = slab + g.final_off;   // every stage already ran (wide first stage only)
        for (int i = tid; i < g.n_final; i += kFastThreads) cp_async8(fin + i, src + i);
    }
    for (int s = g.s_begin; s < g.nstages; s++) {
        const TailStage& st = g.st[s];
        const int hist = st.T - 1;
        if (s > g.s_begin) {
            const float2* __restrict__ hsrc = slab + st.in_off - hist;
            const FastRegion rs = reg[s];
            for (int i = tid; i < hist; i += kFastThreads) cp_async8(&fast_at(x, rs, i), hsrc + i);
        }
        // polyphase bank as it is; FIR taps raw, behind the place of their table
        const int nraw = st.type == TAIL_POLY ? st.interp * st.T : st.T;
        float* rdst = taps + toff[s] + (st.type == TAIL_POLY ? 0 : fast_table_floats(st));
        for (int i = tid; i < nraw; i += kFastThreads) cp_async4(rdst + i, st.taps + i);
    }
    for (int s = g.s_begin; s < g.nstages; s++) {
        const FastRegion rs = reg[s];
        if (rs.M == 1) continue;
        const int D = g.st[s].type == TAIL_DECFIR ? g.st[s].D : 1;
        for (int i = rs.len + tid; i < rs.len + 6 * D + 8; i += kFastThreads) fast_at(x, rs, i) = make_float2(0.0f, 0.0f);
    }
    if (tid == 0) fin[-1] = slab[g.final_off - 1];
    TF_MARK(2);
    cp_async_wait_all();
    __syncthreads();
    for (int s = g.s_begin; s < g.nstages; s++) {
        const TailStage& st = g.st[s];
        if (st.type == TAIL_POLY) continue;
        const int D = st.type == TAIL_DECFIR ? st.D : 1, T = st.T;
        float4* t4 = reinterpret_cast<float4*>(taps + toff[s]);
        const float* __restrict__ raw = taps + toff[s] + fast_table_floats(st);
        for (int sv = tid; sv < T + (kFastOB - 1) * D; sv += kFastThreads) {
            float4 h;
            h.x = sv < T ? raw[sv] : 0.0f;
            h.y = (sv - D >= 0 && sv - D < T) ? raw[sv - D] : 0.0f;
            h.z = (sv - 2 * D >= 0 && sv - 2 * D < T) ? raw[sv - 2 * D] : 0.0f;
            h.w = (sv - 3 * D >= 0 && sv - 3 * D < T) ? raw[sv - 3 * D] : 0.0f;
            t4[sv] = h;
        }
    }
    __syncthreads();
    TF_MARK(3);

    // ---- the stages, back to back out of shared memory ------------------------------------------------------------------
    for (int s = g.s_begin; s < g.nstages; s++) {
        const TailStage& st = g.st[s];
        const int T = st.T;
        const FastRegion ri = reg[s];
        const FastRegion ro = reg[s + 1];
        const int obase = (s + 1 < g.nstages) ? (g.st[s + 1].T - 1) : 1;   // outputs land behind the consumer's history
        if (st.type == TAIL_POLY) {
            const float2* __restrict__ in = x + ri.base;
            const float* __restrict__ h0 = taps + toff[s];
            const int p_phase = st.phase, p_D = st.D, p_interp = st.interp, p_off = st.offset, p_nout = st.n_out;
            for (int o = tid; o < p_nout; o += kFastThreads) {
                // closed form of polyphase_resampler.h:75-93 (phase + o*D fits 32 bits: o < 2^20, D < 2^11)
                const int P = p_phase + o * p_D;
                const int q = P / p_interp;
                const float2* __restrict__ xp = in + p_off + q;
                const float* __restrict__ h = h0 + (P - q * p_interp) * T;
                float2 a0 = make_float2(0.0f, 0.0f), a1 = a0, a2 = a0, a3 = a0;
                int k = 0;
                for (; k + 4 <= T; k += 4) {
                    const float t0 = h[k], t1 = h[k + 1], t2 = h[k + 2], t3 = h[k + 3];
                    a0 = __ffma2_rn(make_float2(t0, t0), xp[k], a0);
                    a1 = __ffma2_rn(make_float2(t1, t1), xp[k + 1], a1);
                    a2 = __ffma2_rn(make_float2(t2, t2), xp[k + 2], a2);
                    a3 = __ffma2_rn(make_float2(t3, t3), xp[k + 3], a3);
                }
                for (; k < T; k++) { const float t = h[k]; a0 = __ffma2_rn(make_float2(t, t), xp[k], a0); }
                fast_at(x, ro, obase + o) = make_float2((a0.x + a1.x) + (a2.x + a3.x), (a0.y + a1.y) + (a2.y + a3.y));
            }
            __syncthreads();
            TF_MARK(4 + s);
            continue;
        }
        // FIR / decimating FIR, register-blocked: thread (og, ks) owns outputs 4*og .. 4*og+3 over a slice of the S = 3D + T
        // window steps; step j meets element offset + 4*og*D + j and the taps (h[j], h[j-D], h[j-2D], h[j-3D]).
        const int D = st.type == TAIL_DECFIR ? st.D : 1;
        const int n_out = st.n_out, st_offset = st.offset;
        const int r_base = ri.base, r_mask = ri.M - 1, r_qs = ri.qs, r_lg = ri.lgM;   // plain registers for the inner loop
        const int G = (n_out + kFastOB - 1) / kFastOB;
        const int S = (kFastOB - 1) * D + T;
        int KS = 1;
        // split the window over as many thread groups as fit in the CTA, down to 8 window steps per thread: the stages are
        // latency chains (~110-220 cycles per window step with a handful of warps busy), so halving the steps halves a stage
        while (KS < 16 && G * KS * 2 <= kFastThreads && S >= 16 * KS) KS *= 2;
        const int Sk = (S + KS - 1) / KS;
        const float4* __restrict__ t4 = reinterpret_cast<const float4*>(taps + toff[s]);
        for (int g0 = 0; g0 < G; g0 += kFastThreads / KS) {
            const int gcount = min(kFastThreads / KS, G - g0);
            const int og = g0 + tid % (kFastThreads / KS), ks = tid / (kFastThreads / KS);
            const bool active = (tid % (kFastThreads / KS)) < gcount && ks < KS;
            float2 acc[kFastOB];
#pragma unroll
            for (int r = 0; r < kFastOB; r++) acc[r] = make_float2(0.0f, 0.0f);
            if (s == g.nstages - 1) TF_MARK(13);
            if (active) {
                const int e0 = st_offset + kFastOB * og * D;     // first element of the group's window
                const int j1 = min(S, (ks + 1) * Sk);
                // element e0 + j sits at plane (e0 + j) % M, position (e0 + j) / M; e0 = offset (mod M) for every group
                const float2* __restrict__ xb = x + r_base;
                int j = ks * Sk;
                // four window steps per trip with all eight loads first (see tail_fir_blocked), then the odd steps
                for (; j + 4 <= j1; j += 4) {
                    const int e = e0 + j;
                    const float2 v0 = xb[(e & r_mask) * r_qs + (e >> r_lg)];
                    const float2 v1 = xb[((e + 1) & r_mask) * r_qs + ((e + 1) >> r_lg)];
                    const float2 v2 = xb[((e + 2) & r_mask) * r_qs + ((e + 2) >> r_lg)];
                    const float2 v3 = xb[((e + 3) & r_mask) * r_qs + ((e + 3) >> r_lg)];
                    const float4 h0 = t4[j], h1 = t4[j + 1], h2 = t4[j + 2], h3 = t4[j + 3];
                    acc[0] = __ffma2_rn(make_float2(h0.x, h0.x), v0, acc[0]);
                    acc[1] = __ffma2_rn(make_float2(h0.y, h0.y), v0, acc[1]);
                    acc[2] = __ffma2_rn(make_float2(h0.z, h0.z), v0, acc[2]);
                    acc[3] = __ffma2_rn(make_float2(h0.w, h0.w), v0, acc[3]);
                    acc[0] = __ffma2_rn(make_float2(h1.x, h1.x), v1, acc[0]);
                    acc[1] = __ffma2_rn(make_float2(h1.y, h1.y), v1, acc[1]);
                    acc[2] = __ffma2_rn(make_float2(h1.z, h1.z), v1, acc[2]);
                    acc[3] = __ffma2_rn(make_float2(h1.w, h1.w), v1, acc[3]);
                    acc[0] = __ffma2_rn(make_float2(h2.x, h2.x), v2, acc[0]);
                    acc[1] = __ffma2_rn(make_float2(h2.y, h2.y), v2, acc[1]);
                    acc[2] = __ffma2_rn(make_float2(h2.z, h2.z), v2, acc[2]);
                    acc[3] = __ffma2_rn(make_float2(h2.w, h2.w), v2, acc[3]);
                    acc[0] = __ffma2_rn(make_float2(h3.x, h3.x), v3, acc[0]);
                    acc[1] = __ffma2_rn(make_float2(h3.y, h3.y), v3, acc[1]);
                    acc[2] = __ffma2_rn(make_float2(h3.z, h3.z), v3, acc[2]);
                    acc[3] = __ffma2_rn(make_float2(h3.w, h3.w), v3, acc[3]);
                }
                for (; j < j1; j++) {
                    const int e = e0 + j;
                    const float2 v = xb[(e & r_mask) * r_qs + (e >> r_lg)];
                    const float4 h = t4[j];
                    acc[0] = __ffma2_rn(make_float2(h.x, h.x), v, acc[0]);
                    acc[1] = __ffma2_rn(make_float2(h.y, h.y), v, acc[1]);
                    acc[2] = __ffma2_rn(make_float2(h.z, h.z), v, acc[2]);
                    acc[3] = __ffma2_rn(make_float2(h.w, h.w), v, acc[3]);
                }
            }
            if (s == g.nstages - 1) TF_MARK(14);
            if (KS > 1) {
                if (active && ks > 0) {
#pragma unroll
                    for (int r = 0; r < kFastOB; r++) scratch[((ks - 1) * (kFastThreads / KS) + (og - g0)) * kFastOB + r] = acc[r];
                }
                __syncthreads();
                if (active && ks == 0) {
                    for (int q = 1; q < KS; q++) {
#pragma unroll
                        for (int r = 0; r < kFastOB; r++) {
                            const float2 v = scratch[((q - 1) * (kFastThreads / KS) + (og - g0)) * kFastOB + r];
                            acc[r].x += v.x; acc[r].y += v.y;
                        }
                    }
                }
            }
            if (active && ks == 0) {
#pragma unroll
                for (int r = 0; r < kFastOB; r++) {
                    const int o = kFastOB * og + r;
                    if (o < n_out) fast_at(x, ro, obase + o) = acc[r];
                }
            }
            if (s == g.nstages - 1) TF_MARK(15);
            if (KS > 1) __syncthreads();   // scratch is reused by the next round of groups
        }
        __syncthreads();
        TF_MARK(4 + s);
    }

    // ---- results, new histories -----------------------------------------------------------------------------------------
    float2* o_iq = a.arena_iq + vd.out_off;
    float* o_dm = a.arena_demod + vd.out_off;
    for (int i = tid; i < g.n_final; i += kFastThreads) {
        const float2 y = fin[i];
        o_iq[i] = y;
        if (g.demod != 0) o_dm[i] = demod_front_end(g, vd, y, fin[i - 1], i);
    }
    for (int s = g.s_begin; s < g.nstages; s++) {
        const TailStage& st = g.st[s];
        const int hist = st.T - 1;
        // stage 0 reads the double-buffered stage-1 region: its history goes to the OTHER region (always, even for an
        // empty block); later stages keep theirs in front of their own input area (fir.h:80)
        if (s == 0 || st.n_in > 0) {
            float2* dst = (s == 0) ? (slab + g.carry0_off - hist) : (slab + st.in_off - hist);
            const FastRegion rs = reg[s];
            const int n_in = st.n_in;
            for (int i = tid; i < hist; i += kFastThreads) dst[i] = fast_at(x, rs, n_in + i);
        }
    }
    if (tid == 0 && g.n_final > 0) slab[g.final_off - 1] = fin[g.n_final - 1];
    TF_MARK(12);
}

cudaError_t launch_tail_fast(Launcher& L, int sid, const TailArgs& a, const TailArgs* d_a, int total_vfos, int threads) {
    if (total_vfos <= 0) return cudaSuccess;
    if (threads != kFastThreadsWide && threads != kFastThreadsNarrow) return cudaErrorInvalidValue;
    size_t smem = 0;
    for (int i = 0; i < a.ngroups; i++) {
        int n = 0, tp = 0;
        if (!tail_fast_fits(a.g[i], &n, threads, &tp)) return cudaErrorInvalidValue;
        smem = std::max(smem, (size_t)fast_tap_bytes(tp) + (size_t)(n + 8) * sizeof(float2));
    }
    const void* fn = threads == kFastThreadsWide ? (const void*)tail_fast_kernel<kFastThreadsWide> : (const void*)tail_fast_kernel<kFastThreadsNarrow>;
    if (cudaError_t e = ensure_dynamic_smem(fn, kFastTapFloats * sizeof(float) + (fast_max_samples(threads) + 8) * sizeof(float2)); e != cudaSuccess) return e;
    return L.kernel(sid, fn, dim3((unsigned)total_vfos), dim3((unsigned)threads), smem, d_a);
}

// ---------------------------------------------------------------------------------------------
// The first tail stage (the second decimating FIR of the PowerDecimator cascade) on its own, wide grid: it still
// runs at 1/D of the input rate for every VFO (e.g. 9600 samples per VFO and block), which one CTA per VFO can only
// walk in three sequential load-compute rounds. Here a CTA of 128 threads owns a range of outputs of one VFO, so the
// stage is spread over the whole machine; the tail kernel then starts from the second stage. CTA 0 of a VFO also
// carries the stage's history to the other stage-1 region (fir.h:80), like the tail kernel does for its stage 0.
// ---------------------------------------------------------------------------------------------
constexpr int kWideThreads = 128;
constexpr int kWideR = 2;                        // outputs per thread: tid + 128 r
constexpr int kWideOut = kWideR * kWideThreads;  // outputs per CTA at most
constexpr int kWideMaxTaps = 128;

// Samples of the CTA's window are staged transposed by D (element i at plane i % D, position i / D, odd plane
// stride): for every tap the threads of a warp (consecutive outputs) read consecutive positions of one plane.
// A tap load serves the thread's four outputs; accumulation is packed FMA on (re, im).
template <int D>
__global__ void __launch_bounds__(kWideThreads, 10)
tail_stage0_wide_kernel(const TailArgs* __restrict__ ap) {
    const TailArgs& a = *ap;   // per-block arguments in the launcher's descriptor (device memory)
    extern __shared__ __align__(16) unsigned char tail_smem[];
    float* taps = reinterpret_cast<float*>(tail_smem);                            // [kWideMaxTaps + 16], zero padded
    float2* xs = reinterpret_cast<float2*>(tail_smem + (kWideMaxTaps + 16) * 4);  // transposed window
    int vi = blockIdx.y, gi = 0;
    while (gi < a.ngroups - 1 && vi >= a.g[gi].nvfo) { vi -= a.g[gi].nvfo; gi++; }
    const TailGroup& g = a.g[gi];
    if (g.s_begin == 0 || g.nstages == 0 || g.st[0].D != D) return; // not this instantiation's (or the tail kernel's) stage
    const TailStage& st = g.st[0];
    float2* slab = a.vfos[g.first_vfo + vi].slab;
    constexpr int lg = D == 2 ? 1 : D == 4 ? 2 : D == 8 ? 3 : 4;
    constexpr int qs = (kWideOut + (kWideMaxTaps >> lg) + 2) | 1;
    const int T = st.T, hist = T - 1;
    const float2* __restrict__ buf = slab + st.in_off - hist;
    float2* __restrict__ out = slab + ((g.nstages > 1) ? g.st[1].in_off : g.final_off);
    const int tid = threadIdx.x;
    // equal ranges of outputs over the CTAs of this VFO
    const int nct = max(1, (st.n_out + kWideOut - 1) / kWideOut);
    const int per = (st.n_out + nct - 1) / nct;
    const int o0 = (int)blockIdx.x * per;
    if ((int)blockIdx.x < nct && o0 < st.n_out) {
        const int co = min(per, st.n_out - o0);
        const int n = (co - 1) * D + T;                 // samples of the window that exist
        const float2* __restrict__ src = buf + st.offset + o0 * D;
        for (int i = tid; i < n; i += kWideThreads) cp_async8(xs + (i & (D - 1)) * qs + (i >> lg), src + i);
        // the last tap group is zero padded to D taps: the samples it meets must be finite
        for (int i = n + tid; i < n + D; i += kWideThreads) xs[(i & (D - 1)) * qs + (i >> lg)] = make_float2(0.0f, 0.0f);
        for (int i = tid; i < kWideMaxTaps + 16; i += kWideThreads) taps[i] = i < T ? __ldg(st.taps + i) : 0.0f;
        cp_async_wait_all();
        __syncthreads();
        float2 acc[kWideR];
#pragma unroll
        for (int r = 0; r < kWideR; r++) acc[r] = make_float2(0.0f, 0.0f);
        // positions past the staged window are only read for outputs >= co (not stored)
        for (int k0 = 0; k0 < T; k0 += D) {
            const float2* __restrict__ xp = xs + (k0 >> lg) + tid;
#pragma unroll
            for (int c = 0; c < D; c++) {
                const float h = taps[k0 + c];
#pragma unroll
                for (int r = 0; r < kWideR; r++) acc[r] = __ffma2_rn(make_float2(h, h), xp[c * qs + r * kWideThreads], acc[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < kWideR; r++)
            if (tid + r * kWideThreads < co) out[o0 + tid + r * kWideThreads] = acc[r];
    }
    if (blockIdx.x == 0) {
        float2* dst = slab + g.carry0_off - hist;
        for (int i = tid; i < hist; i += kWideThreads) dst[i] = buf[st.n_in + i];
    }
}

template <int D>
static cudaError_t launch_wide_t(Launcher& L, int sid, const TailArgs* d_a, int total_vfos, int max_out) {
    constexpr int lg = D == 2 ? 1 : D == 4 ? 2 : D == 8 ? 3 : 4;
    constexpr int qs = (kWideOut + (kWideMaxTaps >> lg) + 2) | 1;
    const size_t smem = (size_t)(kWideMaxTaps + 16) * sizeof(float) + (size_t)D * qs * sizeof(float2);
    if (cudaError_t e = ensure_dynamic_smem((const void*)tail_stage0_wide_kernel<D>, smem); e != cudaSuccess) return e;
    dim3 grid((unsigned)std::max(1, ceil_div(max_out, kWideOut)), (unsigned)total_vfos);
    return L.kernel(sid, (const void*)tail_stage0_wide_kernel<D>, grid, dim3(kWideThreads), smem, d_a);
}

bool tail_stage0_wide_supported(int T, int D) { return T <= kWideMaxTaps && (D == 2 || D == 4 || D == 8 || D == 16); }

cudaError_t launch_tail_stage0_wide(Launcher& L, int sid, const TailArgs& a, const TailArgs* d_a, int total_vfos) {
    if (total_vfos <= 0) return cudaSuccess;
    for (int D : { 2, 4, 8, 16 }) {
        int max_out = 0;
        bool any = false;
        for (int i = 0; i < a.ngroups; i++)
            if (a.g[i].s_begin == 1 && a.g[i].nstages > 0 && a.g[i].st[0].D == D) { any = true; max_out = std::max(max_out, a.g[i].st[0].n_out); }
        if (!any) continue;
        cudaError_t e = D == 2 ? launch_wide_t<2>(L, sid, d_a, total_vfos, max_out) : D == 4 ? launch_wide_t<4>(L, sid, d_a, total_vfos, max_out)
                      : D == 8 ? launch_wide_t<8>(L, sid, d_a, total_vfos, max_out) : launch_wide_t<16>(L, sid, d_a, total_vfos, max_out);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_tail(Launcher& L, int sid, const TailArgs& a, const TailArgs* d_a, int total_vfos) {
    if (total_vfos <= 0) return cudaSuccess;
    const size_t smem = (size_t)kTailTapFloats * sizeof(float) + (size_t)(kTailSmemSamples + 512) * sizeof(float2);
    if (cudaError_t e = ensure_dynamic_smem((const void*)tail_kernel, smem); e != cudaSuccess) return e;
    (void)a;
    return L.kernel(sid, (const void*)tail_kernel, dim3((unsigned)total_vfos), dim3(kTailThreads), smem, d_a);
}


// ---------------------------------------------------------------------------------------------
// Post-detector stages. The AGC and the DC blocker are sequential recurrences with data-dependent
// control flow (clip look-ahead, agc.h:106-121): thread 0 runs them with the reference's operation
// order (_rn intrinsics: no contraction), everything else is parallel over the block's samples.
// ---------------------------------------------------------------------------------------------
constexpr int kPostThreads = 128;
constexpr int kPostTapFloats = 2048;

struct AgcCoef { float attack, inv_attack, decay, inv_decay, set_point, max_gain, max_out; };

// dsp::loop::AGC<T>::process (agc.h:87-147), enabled branch. amp_of(i) = |in[i]|; apply(i, gain) writes the output.
template <class AmpFn, class ApplyFn>
__device__ __forceinline__ void agc_run(const AgcCoef& c, float& amp, float& gain, int count, AmpFn amp_of, ApplyFn apply) {
    for (int i = 0; i < count; i++) {
        const float in_amp = amp_of(i);
        if (in_amp != 0.0f) {
            amp = (in_amp > amp) ? __fadd_rn(__fmul_rn(amp, c.inv_attack), __fmul_rn(in_amp, c.attack))
                                 : __fadd_rn(__fmul_rn(amp, c.inv_decay), __fmul_rn(in_amp, c.decay));
            gain = fminf(__fdiv_rn(c.set_point, amp), c.max_gain);
        } else {
            gain = 1.0f;
        }
        if (__fmul_rn(in_amp, gain) > c.max_out) { // clipping ahead: restart from the largest amplitude left in the block
            float max_amp = 0.0f;
            for (int j = i; j < count; j++) { const float v = amp_of(j); if (v > max_amp) max_amp = v; }
            amp = max_amp;
            gain = fminf(__fdiv_rn(c.set_point, amp), c.max_gain);
        }
        apply(i, gain);
    }
}

// FIR<float,float>::process (fir.h:62-83) on buf = [ntaps-1 history | n samples] (history directly in front of
// `work`), result to out; then the last ntaps-1 samples become the history.
__device__ __forceinline__ void post_fir(float* work, int n, const float* __restrict__ staps, int ntaps, float* __restrict__ out) {
    const float* buf = work - (ntaps - 1);
    for (int i = threadIdx.x; i < n; i += kPostThreads) {
        float acc = 0.0f;
        for (int k = 0; k < ntaps; k++) acc = fmaf(buf[i + k], staps[k], acc);
        out[i] = acc;
    }
    __syncthreads();
    const int hist = ntaps - 1;
    float keep[kPostTapFloats / kPostThreads];
    int c = 0;
    for (int j = threadIdx.x; j < hist; j += kPostThreads, c++) keep[c] = buf[n + j];
    __syncthreads();
    c = 0;
    for (int j = threadIdx.x; j < hist; j += kPostThreads, c++) work[j - hist] = keep[c];
}

// dsp::demod::BroadcastFM::process (demod/broadcast_fm.h:147-214) behind the quadrature front end, RDS output off.
// mpx = the discriminator output. Stereo: the 19 kHz pilot is isolated by a complex band-pass FIR (taps::bandPass<complex_t>,
// applied to (mpx, 0)), a PLL (loop/pll.h:66-72 over loop/phase_control_loop.h:58-66: sequential, one thread, the
// reference's operation order) locks a VCO to it, the mpx delayed by the filter's group delay is multiplied twice by
// conj(vco) -- down-converting the 38 kHz L-R subcarrier -- doubled, and L = (L+R) + (L-R), R = (L+R) - (L-R); both then
// go through the 15 kHz low-pass when enabled. Mono: the low-pass on mpx, copied to both channels.
// State (global, per VFO): [0] PLL phase [1] PLL freq .. [16] pilot (2 cap) | vco (2 cap) | pilot-FIR input [Tp-1 hist | cap] |
// delay line [delay | cap] | L [Ta-1 hist | cap] | R [Ta-1 hist | cap].
__device__ __noinline__ void post_wfm(const PostDev& pd, int n, const float* __restrict__ dm, float* __restrict__ out_l, float* __restrict__ out_r,
                                      float* staps) {
    const int tid = threadIdx.x;
    const bool stereo = (pd.mode & 1) != 0, lowpass = (pd.mode & 2) != 0;
    const int Tp = pd.ntaps, Ta = pd.ntaps2, delay = pd.delay, cap = pd.cap;
    float* st = pd.state;
    float2* pil = reinterpret_cast<float2*>(st + 16);   // [cap] pilot filter output
    float2* vco = pil + cap;                            // [cap] PLL output
    float* pbuf = st + 16 + 4 * cap;             // [Tp-1 | cap]
    float* dbuf = pbuf + (Tp - 1) + cap;         // [delay | cap]
    float* lbuf = dbuf + delay + cap;            // [Ta-1 | cap]
    float* rbuf = lbuf + (Ta - 1) + cap;         // [Ta-1 | cap]
    for (int i = tid; i < Ta; i += kPostThreads) staps[i] = pd.taps2[i];
    if (!stereo) {
        // raw MPX to both channels, through alFir when the low-pass is on (broadcast_fm.h:203-209)
        if (lowpass) {
            for (int i = tid; i < n; i += kPostThreads) lbuf[(Ta - 1) + i] = dm[i];
            __syncthreads();
            post_fir(lbuf + (Ta - 1), n, staps, Ta, out_l);
            __syncthreads();
            for (int i = tid; i < n; i += kPostThreads) out_r[i] = out_l[i];
        } else {
            for (int i = tid; i < n; i += kPostThreads) { out_l[i] = dm[i]; out_r[i] = dm[i]; }
        }
        return;
    }
    for (int i = tid; i < n; i += kPostThreads) { const float v = dm[i]; pbuf[(Tp - 1) + i] = v; dbuf[delay + i] = v; }
    __syncthreads();
    // pilot band-pass: FIR<complex_t, complex_t> on (mpx, 0) -- volk_32fc_x2_dot_prod_32fc with a zero imaginary input
    const float2* __restrict__ ptaps = reinterpret_cast<const float2*>(pd.taps);
    for (int i = tid; i < n; i += kPostThreads) {
        float re = 0.0f, im = 0.0f;
        for (int k = 0; k < Tp; k++) { const float v = pbuf[i + k]; const float2 t = __ldg(ptaps + k); re = fmaf(v, t.x, re); im = fmaf(v, t.y, im); }
        pil[i] = make_float2(re, im);
    }
    __syncthreads();
    if (tid == 0) {
        // loop::PLL::process: out = phasor(phase); advance(normalizePhase(in.phase() - phase))
        constexpr float PI = 3.1415926535f;      // FL_M_PI (math/constants.h)
        const float pdelta = __fsub_rn(PI, -PI);
        float phase = st[0], freq = st[1];
        for (int i = 0; i < n; i++) {
            vco[i] = make_float2(cosf(phase), sinf(phase));
            const float2 p = pil[i];
            float err = __fsub_rn(atan2f(p.y, p.x), phase);
            if (err > PI) err = __fsub_rn(err, __fmul_rn(2.0f, PI));
            else if (err <= -PI) err = __fadd_rn(err, __fmul_rn(2.0f, PI));
            freq = __fadd_rn(freq, __fmul_rn(pd.pll_beta, err));
            if (freq > pd.pll_max_freq) freq = pd.pll_max_freq; else if (freq < pd.pll_min_freq) freq = pd.pll_min_freq;
            phase = __fadd_rn(phase, __fadd_rn(freq, __fmul_rn(pd.pll_alpha, err)));
            while (phase > PI) phase = __fsub_rn(phase, pdelta);
            while (phase < -PI) phase = __fadd_rn(phase, pdelta);
        }
        st[0] = phase; st[1] = freq;
    }
    __syncthreads();
    // lmr = Re((d, 0) * conj(vco) * conj(vco)) * 2 (two volk_32fc_x2_multiply_32fc, ComplexToReal, x2); L = d + lmr, R = d - lmr
    for (int i = tid; i < n; i += kPostThreads) {
        const float d = dbuf[i];
        const float2 c = make_float2(vco[i].x, -vco[i].y);
        const float m1r = __fsub_rn(__fmul_rn(d, c.x), __fmul_rn(0.0f, c.y)), m1i = __fadd_rn(__fmul_rn(d, c.y), __fmul_rn(0.0f, c.x));
        const float m2r = __fsub_rn(__fmul_rn(m1r, c.x), __fmul_rn(m1i, c.y));
        const float lmr = __fmul_rn(m2r, 2.0f);
        const float l = __fadd_rn(d, lmr), r = __fsub_rn(d, lmr);
        if (lowpass) { lbuf[(Ta - 1) + i] = l; rbuf[(Ta - 1) + i] = r; } else { out_l[i] = l; out_r[i] = r; }
    }
    __syncthreads();
    if (lowpass) {
        post_fir(lbuf + (Ta - 1), n, staps, Ta, out_l);
        __syncthreads();
        post_fir(rbuf + (Ta - 1), n, staps, Ta, out_r);
        __syncthreads();
    }
    // carry the pilot filter's input history and the delay line (fir.h:80, math/delay.h:52-61)
    {
        float keep[(2048 + kPostThreads - 1) / kPostThreads];
        int c = 0;
        for (int j = tid; j < Tp - 1; j += kPostThreads, c++) keep[c] = pbuf[n + j];
        __syncthreads();
        c = 0;
        for (int j = tid; j < Tp - 1; j += kPostThreads, c++) pbuf[j] = keep[c];
        c = 0;
        for (int j = tid; j < delay; j += kPostThreads, c++) keep[c] = dbuf[n + j];
        __syncthreads();
        c = 0;
        for (int j = tid; j < delay; j += kPostThreads, c++) dbuf[j] = keep[c];
    }
}

__global__ void __launch_bounds__(kPostThreads)
post_kernel(const PostArgs* __restrict__ ap) {
    const PostArgs& a = *ap;   // per-block arguments in the launcher's descriptor (device memory)
    __shared__ float staps[kPostTapFloats];
    int vi = blockIdx.x, gi = 0;
    while (gi < a.ngroups - 1 && vi >= a.g[gi].nvfo) { vi -= a.g[gi].nvfo; gi++; }
    const int n = a.g[gi].n;
    const PostDev pd = a.post[a.g[gi].first_vfo + vi];
    if (pd.kind == POST_NONE || n <= 0) return;
    const int tid = threadIdx.x;
    const float2* __restrict__ iq = a.arena_iq + pd.out_off;
    const float* __restrict__ dm = a.arena_demod + pd.out_off;
    float* __restrict__ out = a.arena_audio + pd.out_off;
    float* st = pd.state;
    float* work = st + 16 + pd.hist_pad;
    const AgcCoef c{ pd.attack, pd.inv_attack, pd.decay, pd.inv_decay, pd.set_point, pd.max_gain, pd.max_out };
    for (int i = tid; i < pd.ntaps; i += kPostThreads) staps[i] = pd.taps[i];

    if (pd.kind == POST_WFM) {
        post_wfm(pd, n, dm, out, a.arena_audio_r + pd.out_off, staps);
    } else if (pd.kind == POST_FM) {
        if (pd.ntaps > 0) {
            for (int i = tid; i < n; i += kPostThreads) work[i] = dm[i];
            __syncthreads();
            post_fir(work, n, staps, pd.ntaps, out);
        } else {
            for (int i = tid; i < n; i += kPostThreads) out[i] = dm[i];
        }
    } else if (pd.kind == POST_AM) {
        if (pd.mode == 1) {
            // carrier AGC on the complex samples (always enabled, am.h:38), then the magnitude of the scaled sample
            if (tid == 0) {
                float amp = st[2], gain = st[3];
                agc_run(c, amp, gain, n,
                        [&](int i) { const float2 v = iq[i]; return __fsqrt_rn(__fadd_rn(__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y))); },
                        [&](int i, float g) {
                            const float2 v = iq[i];
                            const float re = __fmul_rn(v.x, g), im = __fmul_rn(v.y, g);
                            work[i] = __fsqrt_rn(__fadd_rn(__fmul_rn(re, re), __fmul_rn(im, im)));
                        });
                st[2] = amp; st[3] = gain;
            }
        } else {
            for (int i = tid; i < n; i += kPostThreads) work[i] = dm[i];
        }
        __syncthreads();
        if (tid == 0) {
            // DCBlocker<float> (dc_blocker.h:54-60), then the audio AGC unless the carrier AGC is in use
            float off = st[4];
            for (int i = 0; i < n; i++) {
                const float o = __fadd_rn(work[i], -off);
                work[i] = o;
                off = __fadd_rn(off, __fmul_rn(o, pd.dc_rate));
            }
            st[4] = off;
            if (pd.mode != 1) {
                float amp = st[0], gain = st[1];
                if (pd.mode == 2) {
                    agc_run(c, amp, gain, n, [&](int i) { return fabsf(work[i]); }, [&](int i, float g) { work[i] = __fmul_rn(work[i], g); });
                } else {
                    // disabled AGC: fixed gain with clipping to max_out (agc.h:126-143)
                    for (int i = 0; i < n; i++) {
                        const float v = work[i], in_amp = fabsf(v);
                        work[i] = (__fmul_rn(in_amp, gain) > c.max_out) ? __fmul_rn(v, __fdiv_rn(c.max_out, in_amp)) : __fmul_rn(v, gain);
                    }
                }
                st[0] = amp; st[1] = gain;
            }
        }
        __syncthreads();
        post_fir(work, n, staps, pd.ntaps, out);
    } else {
        if (tid == 0) {
            float amp = st[0], gain = st[1];
            if (pd.mode == 1) {
                agc_run(c, amp, gain, n, [&](int i) { return fabsf(dm[i]); }, [&](int i, float g) { out[i] = __fmul_rn(dm[i], g); });
            } else {
                for (int i = 0; i < n; i++) {
                    const float v = dm[i], in_amp = fabsf(v);
                    out[i] = (__fmul_rn(in_amp, gain) > c.max_out) ? __fmul_rn(v, __fdiv_rn(c.max_out, in_amp)) : __fmul_rn(v, gain);
                }
            }
            st[0] = amp; st[1] = gain;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// RDS side output of BroadcastFM (kernels.h: RdsDev / RdsBlk). Stages run back to back on the VFO's own history buffers;
// a block of a 250 kS/s WFM channel is 1250 samples in and 25 out, so one CTA per VFO and plain loops.
// ---------------------------------------------------------------------------------------------
constexpr int kRdsThreads = 256;

// the last `hist` samples of [hist | n] become the history of the next block (fir.h:80); ascending chunks never read a
// position an earlier chunk has already written
__device__ __forceinline__ void rds_carry(float2* buf, int hist, int n) {
    for (int base = 0; base < hist; base += kRdsThreads) {
        const int j = base + (int)threadIdx.x;
        float2 v = make_float2(0.0f, 0.0f);
        if (j < hist) v = buf[n + j];
        __syncthreads();
        if (j < hist) buf[j] = v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kRdsThreads)
rds_kernel(const RdsArgs* __restrict__ ap) {
    const RdsArgs& a = *ap;
    const RdsDev d = a.tab[a.first + blockIdx.x];
    const RdsBlk& b = a.blk[blockIdx.x];
    const int n = b.n, tid = threadIdx.x;
    if (n <= 0) return;
    const float* __restrict__ dm = a.arena_demod + d.in_off;
    float2* __restrict__ out = a.arena_rds + d.out_off;
    // RealToComplex (x, 0) and FrequencyXlator: out = in * phasor, volk_32fc_s32fc_x2_rotator's complex multiply
    float2* first = d.nstages > 0 ? d.state + d.buf_off[0] + (d.T[0] - 1) : (d.tpp > 0 ? d.state + d.pbuf_off + (d.tpp - 1) : out);
    for (int i = tid; i < n; i += kRdsThreads) {
        const float2 p = phasor_u64(b.phase0 + (uint64_t)i * d.dphi);
        first[i] = cmul(make_float2(dm[i], 0.0f), p);
    }
    __syncthreads();
    int cur = n;
    for (int s = 0; s < d.nstages; s++) {
        const int T = d.T[s], D = d.D[s], nout = b.nout[s];
        float2* buf = d.state + d.buf_off[s];
        float2* dst = (s + 1 < d.nstages) ? d.state + d.buf_off[s + 1] + (d.T[s + 1] - 1) : (d.tpp > 0 ? d.state + d.pbuf_off + (d.tpp - 1) : out);
        const float* __restrict__ h = d.taps[s];
        for (int j = tid; j < nout; j += kRdsThreads) {
            const float2* x = buf + b.off[s] + j * D;       // decimating_fir.h:51-62: dot(&buffer[offset], taps), offset += D
            float2 acc = make_float2(0.0f, 0.0f);
            for (int k = 0; k < T; k++) { const float t = __ldg(h + k); acc = __ffma2_rn(make_float2(t, t), x[k], acc); }
            dst[j] = acc;
        }
        __syncthreads();
        rds_carry(buf, T - 1, cur);
        cur = nout;
    }
    if (d.tpp > 0) {
        float2* buf = d.state + d.pbuf_off;
        for (int m = tid; m < b.npoly; m += kRdsThreads) {
            // closed form of polyphase_resampler.h:75-93
            const long long P = (long long)b.pphase + (long long)m * d.decim;
            const float2* x = buf + b.poff + (int)(P / d.interp);
            const float* __restrict__ h = d.bank + (size_t)(P % d.interp) * d.tpp;
            float2 acc = make_float2(0.0f, 0.0f);
            for (int k = 0; k < d.tpp; k++) { const float t = __ldg(h + k); acc = __ffma2_rn(make_float2(t, t), x[k], acc); }
            out[m] = acc;
        }
        __syncthreads();
        rds_carry(buf, d.tpp - 1, cur);
    }
}

cudaError_t launch_rds(Launcher& L, int sid, const RdsArgs& a) {
    if (a.count <= 0) return cudaSuccess;
    const RdsArgs* d = L.push(a);
    if (!d) return cudaErrorMemoryAllocation;
    return L.kernel(sid, (const void*)rds_kernel, dim3((unsigned)a.count), dim3(kRdsThreads), 0, d);
}

cudaError_t launch_post(Launcher& L, int sid, const PostArgs& a, int total_vfos) {
    if (total_vfos <= 0) return cudaSuccess;
    const PostArgs* d = L.push(a);
    if (!d) return cudaErrorMemoryAllocation;
    return L.kernel(sid, (const void*)post_kernel, dim3((unsigned)total_vfos), dim3(kPostThreads), 0, d);
}

} // namespace sdrpp

#ifdef SDRPP_TAILFAST_TRACE
extern "C" __attribute__((visibility("default"))) int sdrpp_cuda_debug_tailfast_trace(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, sdrpp::g_tailfast_trace, sizeof(long long) * 16);
}
#endif
#ifdef SDRPP_S1_TRACE
extern "C" __attribute__((visibility("default"))) int sdrpp_cuda_debug_s1_trace(long long* out, int rows) {
    if (rows > sdrpp::kS1TraceCap) rows = sdrpp::kS1TraceCap;
    return (int)cudaMemcpyFromSymbol(out, sdrpp::g_s1_trace, sizeof(long long) * 6 * (size_t)rows);
}
#endif
