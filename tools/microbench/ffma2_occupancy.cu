#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__constant__ float ctaps[4096];
// MODE 0: tap scalar from constant (compiler's choice of LDC/LDCU), 36 accumulators, 6 w
// MODE 1: tap as full float2 register pair (h,h) prepared outside loop (no loads in loop)
// MODE 2: scalar FFMA (2 per acc) with tap from constant
template<int MODE>
__global__ void __launch_bounds__(256, 2) k(float2* out, const float2* in, int iters, int off) {
    float2 acc[6][6];
    float2 w[6];
    for (int r = 0; r < 6; r++) { w[r] = in[threadIdx.x + r]; for (int a = 0; a < 6; a++) acc[r][a] = make_float2(r, a); }
    float2 hh[6];
    for (int a = 0; a < 6; a++) hh[a] = in[64 + a];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int a = 0; a < 6; a++) {
            if (MODE == 0) {
                const float h = ctaps[off + it * 6 + a];
#pragma unroll
                for (int r = 0; r < 6; r++) acc[r][a] = __ffma2_rn(make_float2(h, h), w[r], acc[r][a]);
            } else if (MODE == 1) {
#pragma unroll
                for (int r = 0; r < 6; r++) acc[r][a] = __ffma2_rn(hh[a], w[r], acc[r][a]);
            } else {
                const float h = ctaps[off + it * 6 + a];
#pragma unroll
                for (int r = 0; r < 6; r++) { acc[r][a].x = fmaf(h, w[r].x, acc[r][a].x); acc[r][a].y = fmaf(h, w[r].y, acc[r][a].y); }
            }
        }
    }
    float2 s = make_float2(0, 0);
    for (int r = 0; r < 6; r++) for (int a = 0; a < 6; a++) { s.x += acc[r][a].x; s.y += acc[r][a].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template<int MODE> void run(float2* out, float2* in, const char* name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 600;
    for (int threads = 64; threads <= 256; threads *= 2)
    for (int bps = 1; bps <= 2; bps++) {
        float best = 1e9;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            k<MODE><<<148 * bps, threads>>>(out, in, iters, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double fma = (double)148 * bps * threads * iters * 36 * 2;
        printf("%-22s warps/SMSP %4.1f : %.4f ms  %.2f TFMA/s (%.0f%% of 36.6)  %s\n", name, bps * threads / 128.0, best, fma / best / 1e9, 100 * fma / best / 1e9 / 36.6, cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    float2* out; float2* in; cudaMalloc(&out, 148*8*256*8); cudaMalloc(&in, 1024*8); cudaMemset(in, 0, 1024*8);
    run<0>(out, in, "FFMA2 tap const");
    run<1>(out, in, "FFMA2 tap reg pair");
    run<2>(out, in, "FFMA tap const");
    return 0;
}
