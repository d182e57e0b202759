// tcgen05.mma issue/execute rate on one SM per N (M = 128, K = 16, fp16 -> fp32, both operands in shared memory,
// K-major SWIZZLE_128B): cycles per MMA for N = 32..256, with the A/B descriptors (a) fixed and (b) walking over a
// 96 KB + 128 KB operand area like the channelizer's stage-1 kernel does, alone and with the stage-1 kernel's other
// traffic running beside it: tcgen05.ld of the second accumulator stage, bulk copies into shared memory.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o umma_rate umma_rate.cu && ./umma_rate
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}

__global__ void __launch_bounds__(256, 1) k(int N, int iters, int walk, int two_acc, int ldtm_warps, int tma, const uint8_t* gsrc, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar, cbar;
    __shared__ volatile int done;
    __shared__ uint32_t tslot;
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 220 * 1024 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&cbar)));
        done = 0;
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tslot)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tm = tslot;
    if (warp == 1) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t sA = smem_addr(base), sB = smem_addr(base + 96 * 1024);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i++) {
            // 4 k-steps of a 16 KB A chunk and of a B k-half, like one product of the stage-1 kernel
            const uint32_t ao = walk ? (uint32_t)(i % 6) * 16384u : 0u;
            const uint32_t bo = walk ? (uint32_t)(i % 4) * (uint32_t)N * 128u : 0u;
            const uint64_t ad = desc_sw128(sA + ao), bd = desc_sw128(sB + bo);
            const uint32_t d = tm + (two_acc ? (uint32_t)(i & 1) * 256u : 0u);
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 4; ks++) mma(d, ad + 2u * ks, bd + 2u * ks, idesc, 1u);
            }
            __syncwarp();
        }
        if (elect_one()) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(&bar)) : "memory");
        __syncwarp();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_addr(&bar)) : "memory");
        const long long t1 = clock64();
        if ((threadIdx.x & 31) == 0) { out[blockIdx.x] = t1 - t0; done = 1; }
    } else if (warp >= 4 && warp < 4 + ldtm_warps) {
        // concurrent TMEM reads (the epilogue's tcgen05.ld of the OTHER accumulator stage)
        uint32_t r[16], acc = 0;
        while (!done) {
            const uint32_t ta = tm + ((uint32_t)((warp & 3) * 32) << 16) + 256u;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                             : "r"(ta + 16u * j) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                acc += r[0] + r[15];
            }
        }
        if (acc == 0x12345678u) out[0] = 0;
    } else if (warp == 2 && tma) {
        // concurrent bulk copies into an unused part of shared memory (the producer's A chunks)
        uint32_t ph = 0;
        while (!done) {
            if (elect_one()) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(&cbar)), "r"(16384u) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_addr(base + 200 * 1024)), "l"(gsrc + (size_t)blockIdx.x * 16384), "r"(16384u), "r"(smem_addr(&cbar)) : "memory");
            }
            __syncwarp();
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_addr(&cbar)), "r"(ph) : "memory");
            ph ^= 1;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u));
}

int main() {
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    uint8_t* gsrc;
    cudaMalloc(&gsrc, 148 * 16384);
    cudaMemset(gsrc, 0, 148 * 16384);
    const size_t smem = 226 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int iters = 2000;
    const int ctas = 148, walk = 1, two = 0;
    for (int tma = 0; tma < 2; tma++)
        for (int lw : { 0, 1, 4 })
                for (int N : { 32, 64, 96, 128, 192, 224, 256 }) {
                    k<<<ctas, 256, smem>>>(N, iters, walk, two, lw, tma, gsrc, d);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                    long long h[148];
                    cudaMemcpy(h, d, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
                    double s = 0;
                    for (int i = 0; i < ctas; i++) s += (double)h[i];
                    s /= ctas;
                    printf("bulk copies %d, tcgen05.ld warps %d, N %3d: %.1f cycles per MMA (floor %d), %.0f FMA/clk/SM\n", tma, lw, N,
                           s / (iters * 4.0), N / 2, 128.0 * N * 16 * iters * 4 / s);
                }
    return 0;
}
