// Register/shared-memory FFT building blocks for the spectrum kernels.
//
// A CTA computes B independent forward DFTs of length L (unnormalised, sign -1, like FFTW_FORWARD,
// signal_path/iq_frontend.cpp:292). T = L/E threads cooperate on one DFT, each holding E samples in
// registers in the layout idx = t + T*e. The transform is a Stockham autosort with up to three
// passes of radix R0,R1,R2 (R0*R1*R2 = L, R0 = E); each radix-R butterfly is a fully unrolled
// register DFT with compile-time twiddles, and passes exchange data through shared memory once.
// The first pass therefore reads straight from global memory and the last pass leaves results in
// registers (layout k = t + T*e), so a 1024-point transform touches shared memory exactly once.
#pragma once
#include "common.cuh"

namespace sdrpp {

// cos(2*pi*m/32), m = 0..8
__device__ __host__ constexpr float cos32_tab(int m) {
    return m == 0 ? 1.0f : m == 1 ? 0.98078528040323043f : m == 2 ? 0.92387953251128674f :
           m == 3 ? 0.83146961230254524f : m == 4 ? 0.70710678118654752f : m == 5 ? 0.55557023301960218f :
           m == 6 ? 0.38268343236508978f : m == 7 ? 0.19509032201612825f : 0.0f;
}
__device__ __host__ constexpr float cos32(int m) {
    m &= 31;
    if (m > 16) m = 32 - m;
    return m > 8 ? -cos32_tab(16 - m) : cos32_tab(m);
}
__device__ __host__ constexpr float sin32(int m) { return cos32(m - 8); }

// v * exp(-2*pi*i*M/32)
template <int M>
__device__ __forceinline__ float2 mul_w32(float2 v) {
    constexpr int m = M & 31;
    if constexpr (m == 0) return v;
    else if constexpr (m == 8) return make_float2(v.y, -v.x);
    else if constexpr (m == 16) return make_float2(-v.x, -v.y);
    else if constexpr (m == 24) return make_float2(-v.y, v.x);
    else {
        constexpr float c = cos32(m), s = sin32(m);
        return make_float2(v.x * c + v.y * s, v.y * c - v.x * s);
    }
}

// complex add / subtract as one packed instruction (FADD2, sm_100): the (re, im) pair of a float2 is the instruction's pair
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }

template <int R> __device__ __forceinline__ void dft_reg(float2 (&v)[R]);

template <int R, int K>
struct DftCombine {
    __device__ __forceinline__ static void run(float2 (&v)[R], const float2 (&e)[R / 2], const float2 (&o)[R / 2]) {
        const float2 t = mul_w32<K * (32 / R)>(o[K]);
        v[K] = cadd(e[K], t);
        v[K + R / 2] = csub(e[K], t);
        if constexpr (K + 1 < R / 2) DftCombine<R, K + 1>::run(v, e, o);
    }
};

// In-place forward DFT of R register values, natural order in and out (R = 1,2,4,8,16,32).
template <int R>
__device__ __forceinline__ void dft_reg(float2 (&v)[R]) {
    if constexpr (R == 2) {
        const float2 a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    } else if constexpr (R > 2) {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; i++) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        dft_reg<R / 2>(e);
        dft_reg<R / 2>(o);
        DftCombine<R, 0>::run(v, e, o);
    }
}

template <int L_, int E_, int R0_, int R1_, int R2_>
struct FftPlan {
    static constexpr int L = L_, E = E_, R0 = R0_, R1 = R1_, R2 = R2_;
    static constexpr int T = L / E;
    static constexpr int PASSES = (R1 > 1 ? (R2 > 1 ? 3 : 2) : 1);
    static_assert(R0 * R1 * R2 == L, "radices must multiply to L");
    static_assert(R0 == E && E % R1 == 0 && E % R2 == 0, "register tile must hold whole butterflies");
    // row layout: one padding element per R0 elements keeps the stride-R0 scatter of pass 0 conflict-free
    static constexpr int LP = L + L / R0;
};

template <class P, bool COLS, int B>
__device__ __forceinline__ int smem_index(int idx, int b) {
    if constexpr (COLS) return (idx + idx / P::R0) * B + b;   // same padding: the pass-0 scatter of threads t, t+1 lands B*(R0+1) apart
    else return b * P::LP + idx + idx / P::R0;
}

// One Stockham pass on the register tile. NS = product of the radices of earlier passes.
// tw: shared-memory table exp(-2*pi*i*j/L), j < L (correctly rounded on the host) -- every twiddle is a
// single table value, so round-off matches a table-driven CPU FFT instead of growing along a product chain.
template <class P, int R, int NS, bool LAST, bool COLS, int B>
__device__ __forceinline__ void fft_pass(float2 (&v)[P::E], float2* sm, const float2* tw, int t, int b) {
    constexpr int E = P::E, T = P::T, S = E / R, L = P::L;
#pragma unroll
    for (int s = 0; s < S; s++) {
        float2 u[R];
#pragma unroll
        for (int r = 0; r < R; r++) u[r] = v[s + r * S];
        const int j = t + T * s;
        if constexpr (NS > 1) {
            const int k = j % NS;
            constexpr int STRIDE = L / (NS * R);
#pragma unroll
            for (int r = 1; r < R; r++) u[r] = cmul(u[r], tw[(k * r * STRIDE) & (L - 1)]);
        }
        dft_reg<R>(u);
        if constexpr (LAST) {
#pragma unroll
            for (int r = 0; r < R; r++) v[s + r * S] = u[r];
        } else {
            const int j0 = (j / NS) * NS * R + (j % NS);
#pragma unroll
            for (int r = 0; r < R; r++) sm[smem_index<P, COLS, B>(j0 + r * NS, b)] = u[r];
        }
    }
    if constexpr (!LAST) {
        // the callers' twiddle tables arrive by cp.async (fft.cu: stage_table); the first exchange publishes them. Waiting here,
        // behind the first butterflies' stores, keeps the wait out of the frame loads (ahead of block_fft the compiler hoisted
        // it into the middle of them: one more serialised round trip)
        if constexpr (NS == 1) asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int e = 0; e < E; e++) v[e] = sm[smem_index<P, COLS, B>(t + T * e, b)];
        __syncthreads();
    }
}

// Full length-L transform of the register tile (in: x[t+T*e], out: X[t+T*e]).
template <class P, bool COLS, int B>
__device__ __forceinline__ void block_fft(float2 (&v)[P::E], float2* sm, const float2* tw, int t, int b) {
    fft_pass<P, P::R0, 1, P::PASSES == 1, COLS, B>(v, sm, tw, t, b);
    if constexpr (P::PASSES >= 2) fft_pass<P, P::R1, P::R0, P::PASSES == 2, COLS, B>(v, sm, tw, t, b);
    if constexpr (P::PASSES >= 3) fft_pass<P, P::R2, P::R0 * P::R1, true, COLS, B>(v, sm, tw, t, b);
}

// exchange buffer elements (float2) for B transforms
template <class P, bool COLS, int B>
constexpr size_t fft_exchange_elems() {
    return P::PASSES == 1 ? 0 : (size_t)P::LP * B;
}

} // namespace sdrpp
