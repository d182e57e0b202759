#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
__device__ __forceinline__ void fma2(float2& d, float2 a, float2 b) {
    uint64_t dd = *reinterpret_cast<uint64_t*>(&d), aa = *reinterpret_cast<uint64_t*>(&a), bb = *reinterpret_cast<uint64_t*>(&b);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
    d = *reinterpret_cast<float2*>(&dd);
}
template<int MODE>
__global__ void __launch_bounds__(256) k(float* out, const float2* in, int iters) {
    float2 acc[16];
    float2 b[4];
    for (int i = 0; i < 16; i++) acc[i] = make_float2(threadIdx.x * 0.001f + i, i);
    for (int i = 0; i < 4; i++) b[i] = in[i + threadIdx.x % 4];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 16; i++) {
                if (MODE == 0) { acc[i].x = fmaf(acc[i].x, b[u].x, b[(u+1)&3].x); acc[i].y = fmaf(acc[i].y, b[u].y, b[(u+1)&3].y); }
                else { float2 t = b[(u+1)&3]; uint64_t dd = *reinterpret_cast<uint64_t*>(&acc[i]), aa=*reinterpret_cast<uint64_t*>(&b[u]), cc=*reinterpret_cast<uint64_t*>(&t);
                       asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(dd) : "l"(aa), "l"(cc)); acc[i] = *reinterpret_cast<float2*>(&dd); }
            }
        }
    }
    float s = 0; for (int i = 0; i < 16; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    float* out; float2* in; cudaMalloc(&out, 148*8*256*4); cudaMalloc(&in, 64*8); cudaMemset(in, 0, 64*8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; mode++) for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<148*4, 256>>>(out, in, iters); else k<1><<<148*4, 256>>>(out, in, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double fma = (double)148*4*256 * iters * 4 * 16 * 2;
        printf("mode %d: %.3f ms, %.2f TFMA/s (%.2f TFLOP/s) err=%s\n", mode, ms, fma/ms/1e9, 2*fma/ms/1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
