#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
// 36 independent FFMA2 (or 72 FFMA) per iteration + NX extra integer adds (alu pipe) per iteration
template<int MODE, int NX>
__global__ void __launch_bounds__(256, 2) k(float2* out, const float2* in, int iters, int seed) {
    float2 acc[36];
    float2 w[6], hh[6];
    for (int r = 0; r < 6; r++) { w[r] = in[threadIdx.x + r]; hh[r] = in[64 + r]; }
    for (int a = 0; a < 36; a++) acc[a] = make_float2(a, a);
    int z[12];
    for (int i = 0; i < 12; i++) z[i] = seed * (i + 1) + threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int a = 0; a < 36; a++) {
            if (MODE == 0) acc[a] = __ffma2_rn(hh[a / 6], w[a % 6], acc[a]);
            else { acc[a].x = fmaf(hh[a / 6].x, w[a % 6].x, acc[a].x); acc[a].y = fmaf(hh[a / 6].y, w[a % 6].y, acc[a].y); }
            if (a < NX) { asm volatile("add.s32 %0, %0, %1;" : "+r"(z[a % 12]) : "r"(seed)); }
        }
    }
    float2 s = make_float2(0, 0);
    for (int a = 0; a < 36; a++) { s.x += acc[a].x; s.y += acc[a].y; }
    int zz = 0; for (int i = 0; i < 12; i++) zz += z[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = make_float2(s.x + zz, s.y);
}
template<int MODE, int NX> void run(float2* out, float2* in, const char* name) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<MODE, NX><<<148 * 2, 256>>>(out, in, iters, 3);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double fma = (double)148 * 2 * 256 * iters * 36 * 2;
    printf("%-34s %.3f ms  %.2f TFMA/s (%.0f%%)  %s\n", name, best, fma / best / 1e9, 100 * fma / best / 1e9 / 37.2, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float2* out; float2* in; cudaMalloc(&out, 148*8*256*8); cudaMalloc(&in, 1024*8); cudaMemset(in, 0, 1024*8);
    run<0, 0>(out, in, "36 FFMA2");
    run<0, 6>(out, in, "36 FFMA2 + 6 IADD");
    run<0, 12>(out, in, "36 FFMA2 + 12 IADD");
    run<0, 18>(out, in, "36 FFMA2 + 18 IADD");
    run<0, 36>(out, in, "36 FFMA2 + 36 IADD");
    run<1, 0>(out, in, "72 FFMA");
    run<1, 12>(out, in, "72 FFMA + 12 IADD");
    run<1, 36>(out, in, "72 FFMA + 36 IADD");
    return 0;
}
