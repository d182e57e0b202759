// Batched multi-VFO channelizer. Replaces, for every VFO at once, the per-VFO thread of the
// reference: Splitter memcpy (dsp/routing/splitter.h:46-60) -> dsp::channel::RxVFO::process
// (dsp/channel/rx_vfo.h:89-100) = FrequencyXlator (frequency_xlator.h:43-50) -> RationalResampler
// (PowerDecimator cascade power_decimator.h:51-67 + PolyphaseResampler polyphase_resampler.h:69-99)
// -> channel FIR (filter/fir.h:62-83) -> demod front end (demod/quadrature.h:41-56, am.h:122,
// ssb.h:90-101).
//
// Stage 1 (stage1_kernel) carries ~95 % of the arithmetic at high decimation: the NCO and the
// first decimating FIR, which both run at the full input rate. The translation is folded into the
// taps: with z[n] = x[n] e^{j phi(n)} and phi(n0+k) = phi(n0) + w k,
//     y[m] = sum_k z[n0+k] h[k] = e^{j phi(n0)} sum_k x[n0+k] g[k],   g[k] = h[k] e^{j w k},
// so every VFO of a group reads the SAME untranslated samples from one shared-memory tile and
// only its taps differ. A warp holds 32 VFOs (one per lane) x 8 consecutive outputs; the sample
// window is a shared-memory broadcast, the per-VFO taps stream through a 3-stage
// cp.async.bulk (TMA) + mbarrier pipeline, and the inner loop is register-blocked FP32 FMA.
//
// The remaining stages run at 1/D of the rate (tail_kernel, one CTA per VFO).
#include "common.cuh"
#include "kernels.h"

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

__device__ __forceinline__ float2 phasor_u64(uint64_t phase) {
    // phase in turns * 2^64 -> (cos, sin). The top 32 bits as a signed fraction of a half turn.
    const float x = (float)(int32_t)(phase >> 32) * 4.656612873077393e-10f; // * 2^-31
    float s, c;
    sincospif(x, &s, &c);
    return make_float2(c, s);
}

// ---------------------------------------------------------------------------------------------
// Stage 1
// ---------------------------------------------------------------------------------------------
int stage1_pcp(int D) { return (D / 2) < 4 ? (D / 2) : 4; }
bool stage1_supported(int A, int D) { return D >= 2 && (D & (D - 1)) == 0 && D <= 128 && A >= 1 && A <= 8; }

size_t stage1_g_elems(int A, int D, int nvfo) {
    const int nvb = ceil_div(nvfo, 32);
    return (size_t)nvb * (size_t)(D / 2) * (size_t)A * 32;
}
void stage1_g_index(int A, int D, int pcp, int v, int k, size_t* idx4, int* half) {
    const int vb = v / 32, lane = v % 32;
    const int a = k / D, p = k % D, pp = p / 2;
    const int nch = (D / 2) / pcp;
    const int c = pp / pcp, ppc = pp % pcp;
    *idx4 = ((((size_t)vb * nch + c) * A + a) * pcp + ppc) * 32 + lane;
    *half = p & 1;
}

template <int A>
__global__ void __launch_bounds__(kStage1Warps * 32, 2)
stage1_kernel(const __grid_constant__ Stage1Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int MT = kStage1MT, TM = kStage1TM, S = kStage1Stages;
    constexpr int QP = TM + A - 1; // sample rows (of D samples) needed by TM outputs
    constexpr int QS = QP | 1;     // odd row stride in float4 units: conflict-free transposed stores
    const int D = a.D, DP = D >> 1;
    const int pcp = a.pcp, nch = DP / pcp;
    const int chunk_elems = A * pcp * 32;
    const uint32_t chunk_bytes = (uint32_t)chunk_elems * 16u;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
    float4* gs = reinterpret_cast<float4*>(smem_raw + 128);
    float4* xs4 = gs + S * chunk_elems;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m0 = blockIdx.x * TM;
    const int vb = blockIdx.y;
    const float4* __restrict__ Gvb = a.G + (size_t)vb * nch * chunk_elems;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; s++) mbar_init(&bars[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int c = 0; c < S - 1 && c < nch; c++) {
            mbar_expect_tx(&bars[c], chunk_bytes);
            bulk_g2s(gs + c * chunk_elems, Gvb + (size_t)c * chunk_elems, chunk_bytes, &bars[c]);
        }
    }

    // Sample tile, transposed to [pair p/2][row q] so that the window of one tap phase is contiguous.
    {
        const uint32_t r0 = a.ring_first + (uint32_t)m0 * (uint32_t)D;
        const int64_t abs0 = a.abs_first + (int64_t)m0 * D;
        const int total = QP * DP;
        for (int idx = tid; idx < total; idx += kStage1Warps * 32) {
            const int q = idx / DP, pp = idx - q * DP;
            const int n = q * D + 2 * pp;
            float2 s0 = a.ring.base[(r0 + (uint32_t)n) & a.ring.mask];
            float2 s1 = a.ring.base[(r0 + (uint32_t)n + 1u) & a.ring.mask];
            if (abs0 + n < a.abs_valid) s0 = make_float2(0.0f, 0.0f);
            if (abs0 + n + 1 < a.abs_valid) s1 = make_float2(0.0f, 0.0f);
            xs4[pp * QS + q] = make_float4(s0.x, s0.y, s1.x, s1.y);
        }
    }

    float2 acc[MT];
#pragma unroll
    for (int i = 0; i < MT; i++) acc[i] = make_float2(0.0f, 0.0f);
    const int mq = warp * MT;
    const int rem = a.T - (A - 1) * D; // taps in the last row of the tap matrix

    for (int c = 0; c < nch; c++) {
        const int s = c % S;
        mbar_wait(&bars[s], (uint32_t)((c / S) & 1));
        __syncthreads(); // every warp is done with chunk c-1 (and, for c == 0, the sample tile is complete)
        if (tid == 0 && c + S - 1 < nch) {
            const int cn = c + S - 1, sn = cn % S;
            fence_proxy_async();
            mbar_expect_tx(&bars[sn], chunk_bytes);
            bulk_g2s(gs + sn * chunk_elems, Gvb + (size_t)cn * chunk_elems, chunk_bytes, &bars[sn]);
        }
        const float4* g = gs + s * chunk_elems + lane;
        for (int ppc = 0; ppc < pcp; ppc++) {
            const int pp = c * pcp + ppc;
            float4 xw[MT + A - 1];
            const float4* xrow = xs4 + pp * QS + mq;
#pragma unroll
            for (int i = 0; i < MT + A - 1; i++) xw[i] = xrow[i];
#pragma unroll
            for (int aa = 0; aa < A; aa++) {
                if (aa == A - 1 && 2 * pp >= rem) continue; // this slab holds only zero padding
                const float4 g4 = g[(aa * pcp + ppc) * 32];
#pragma unroll
                for (int i = 0; i < MT; i++) {
                    const float4 x = xw[aa + i];
                    acc[i].x = fmaf(x.x, g4.x, acc[i].x);
                    acc[i].x = fmaf(-x.y, g4.y, acc[i].x);
                    acc[i].y = fmaf(x.x, g4.y, acc[i].y);
                    acc[i].y = fmaf(x.y, g4.x, acc[i].y);
                    acc[i].x = fmaf(x.z, g4.z, acc[i].x);
                    acc[i].x = fmaf(-x.w, g4.w, acc[i].x);
                    acc[i].y = fmaf(x.z, g4.w, acc[i].y);
                    acc[i].y = fmaf(x.w, g4.z, acc[i].y);
                }
            }
        }
    }

    const int v = vb * 32 + lane;
    if (v < a.nvfo) {
        const VfoDev vd = a.vfos[v];
        float2* __restrict__ out = vd.slab + a.out_off;
#pragma unroll
        for (int i = 0; i < MT; i++) {
            const int m = m0 + mq + i;
            if (m < a.M) {
                const uint64_t ph = vd.phi_ref + (uint64_t)(a.abs_first + (int64_t)m * D - vd.n_ref) * vd.dphi;
                out[m] = cmul(acc[i], phasor_u64(ph));
            }
        }
    }
}

template <int A>
static cudaError_t launch_stage1_t(const Stage1Args& a, cudaStream_t st) {
    constexpr int QS = (kStage1TM + A - 1) | 1;
    const size_t smem = 128 + (size_t)kStage1Stages * A * a.pcp * 32 * 16 + (size_t)(a.D / 2) * QS * 16;
    static size_t attr_set = 0;
    if (smem > attr_set) {
        cudaError_t e = cudaFuncSetAttribute(stage1_kernel<A>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = smem;
    }
    dim3 grid(ceil_div(a.M, kStage1TM), ceil_div(a.nvfo, 32));
    stage1_kernel<A><<<grid, kStage1Warps * 32, smem, st>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_stage1(const Stage1Args& a, cudaStream_t st) {
    if (a.M <= 0 || a.nvfo <= 0) return cudaSuccess;
    switch (a.A) {
    case 1: return launch_stage1_t<1>(a, st);
    case 2: return launch_stage1_t<2>(a, st);
    case 3: return launch_stage1_t<3>(a, st);
    case 4: return launch_stage1_t<4>(a, st);
    case 5: return launch_stage1_t<5>(a, st);
    case 6: return launch_stage1_t<6>(a, st);
    case 7: return launch_stage1_t<7>(a, st);
    case 8: return launch_stage1_t<8>(a, st);
    }
    return cudaErrorInvalidValue;
}

// D = 1: translation only (RationalResampler modes RESAMP_ONLY / NONE, rational_resampler.h:83-97)
__global__ void __launch_bounds__(256)
mix_only_kernel(const __grid_constant__ Stage1Args a) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    const int v = blockIdx.y;
    if (m >= a.M) return;
    const VfoDev vd = a.vfos[v];
    float2 x = a.ring.base[(a.ring_first + (uint32_t)m) & a.ring.mask];
    if (a.abs_first + m < a.abs_valid) x = make_float2(0.0f, 0.0f);
    const uint64_t ph = vd.phi_ref + (uint64_t)(a.abs_first + (int64_t)m - vd.n_ref) * vd.dphi;
    vd.slab[a.out_off + m] = cmul(x, phasor_u64(ph));
}

cudaError_t launch_mix_only(const Stage1Args& a, cudaStream_t st) {
    if (a.M <= 0 || a.nvfo <= 0) return cudaSuccess;
    dim3 grid(ceil_div(a.M, 256), a.nvfo);
    mix_only_kernel<<<grid, 256, 0, st>>>(a);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Tail: remaining decimating FIRs, polyphase resampler, channel FIR, demod front end.
// One CTA per VFO; stages run back to back on the VFO's slab (L2-resident), __syncthreads between.
// ---------------------------------------------------------------------------------------------
constexpr int kTailThreads = 256;

__global__ void __launch_bounds__(kTailThreads)
tail_kernel(const __grid_constant__ TailArgs a) {
    int vi = blockIdx.x, gi = 0;
    while (gi < a.ngroups - 1 && vi >= a.g[gi].nvfo) { vi -= a.g[gi].nvfo; gi++; }
    const TailGroup& g = a.g[gi];
    const VfoDev vd = a.vfos[g.first_vfo + vi];
    float2* slab = vd.slab;
    const int tid = threadIdx.x;

    for (int s = 0; s < g.nstages; s++) {
        const TailStage& st = g.st[s];
        const int hist = st.T - 1;
        float2* buf = slab + st.in_off - hist; // [hist | n_in]
        float2* out = slab + ((s + 1 < g.nstages) ? g.st[s + 1].in_off : g.final_off);
        for (int m = tid; m < st.n_out; m += kTailThreads) {
            int off;
            const float* taps = st.taps;
            if (st.type == TAIL_POLY) {
                const long long P = (long long)st.phase + (long long)m * st.D;
                off = st.offset + (int)(P / st.interp);
                taps += (size_t)(P % st.interp) * st.T;
            } else {
                off = st.offset + m * st.D;
            }
            const float2* x = buf + off;
            float re = 0.0f, im = 0.0f;
            for (int k = 0; k < st.T; k++) {
                const float2 v = x[k];
                const float h = __ldg(taps + k);
                re = fmaf(v.x, h, re);
                im = fmaf(v.y, h, im);
            }
            out[m] = make_float2(re, im);
        }
        __syncthreads();
        // carry the last `hist` inputs to the front (fir.h:80, decimating_fir.h:65, polyphase_resampler.h:96)
        if (st.n_in > 0) {
            float2 keep[8];
            int n = 0;
            for (int i = tid; i < hist && n < 8; i += kTailThreads, n++) keep[n] = buf[st.n_in + i];
            __syncthreads();
            n = 0;
            for (int i = tid; i < hist && n < 8; i += kTailThreads, n++) buf[i] = keep[n];
            __syncthreads();
        }
    }

    // final output, demod front end, results arena
    float2* fin = slab + g.final_off; // fin[-1] = last output of the previous block
    float2* o_iq = a.arena_iq + vd.out_off;
    float* o_dm = a.arena_demod + vd.out_off;
    for (int i = tid; i < g.n_final; i += kTailThreads) {
        const float2 y = fin[i];
        o_iq[i] = y;
        if (g.demod == 1) {
            const float2 p = fin[i - 1];
            // y * conj(prev) with complex_t::operator* (dsp/types.h:23-25)
            const float dre = y.x * p.x + y.y * p.y;
            const float dim = y.y * p.x - y.x * p.y;
            o_dm[i] = atan2f(dim, dre) * g.inv_dev;
        } else if (g.demod == 2) {
            o_dm[i] = sqrtf(y.x * y.x + y.y * y.y);
        } else if (g.demod >= 3) {
            const float2 w = phasor_u64((uint64_t)(g.abs_out + i) * vd.dphi2);
            o_dm[i] = y.x * w.x - y.y * w.y;
        }
    }
    __syncthreads();
    if (tid == 0 && g.n_final > 0) fin[-1] = fin[g.n_final - 1];
}

cudaError_t launch_tail(const TailArgs& a, int total_vfos, cudaStream_t st) {
    if (total_vfos <= 0) return cudaSuccess;
    tail_kernel<<<total_vfos, kTailThreads, 0, st>>>(a);
    return cudaGetLastError();
}

} // namespace sdrpp
