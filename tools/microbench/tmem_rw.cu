// Micro-benchmark: tensor memory (TMEM) as thread-private storage through tcgen05.ld / tcgen05.st .32x32b (B200).  Not product code.
// 12 warps per CTA, one CTA per SM; warp w owns TMEM lanes 32 (w % 4) .. +31 and columns 96 (w / 4) .. +95.
// Checks that a thread reads back what it stored and prints the read-modify-write rate.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void tm_ld16(uint32_t addr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr));
}
__device__ __forceinline__ void tm_st16(uint32_t addr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
                 "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(384, 1) k(uint32_t* out, int iters, long long* cyc) {
    __shared__ uint32_t tbase;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"l"((uint64_t)__cvta_generic_to_shared(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tbase + ((uint32_t)(32 * (w & 3)) << 16) + 96u * (w >> 2);
    uint32_t r[16];
    for (int c = 0; c < 4; c++) {   // 64 words per thread
        for (int j = 0; j < 16; j++) r[j] = threadIdx.x * 1000u + c * 16 + j;
        tm_st16(base + 16 * c, r);
    }
    tm_wait_st();
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++)
        for (int c = 0; c < 4; c++) {
            tm_ld16(base + 16 * c, r);
            tm_wait_ld();
            for (int j = 0; j < 16; j++) r[j] += 1u;
            tm_st16(base + 16 * c, r);
            tm_wait_st();
        }
    const long long t1 = clock64();
    uint32_t bad = 0;
    for (int c = 0; c < 4; c++) {
        tm_ld16(base + 16 * c, r);
        tm_wait_ld();
        for (int j = 0; j < 16; j++) bad += (r[j] != threadIdx.x * 1000u + c * 16 + j + (uint32_t)iters);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = bad;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}
int main() {
    uint32_t* out; long long* cyc; cudaMalloc(&out, 148 * 384 * 4); cudaMalloc(&cyc, 8);
    const int iters = 1000;
    k<<<148, 384>>>(out, iters, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    static uint32_t h[148 * 384]; long long c = 0;
    cudaMemcpy(h, out, sizeof h, cudaMemcpyDeviceToHost); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    long bad = 0; for (int i = 0; i < 148 * 384; i++) bad += h[i];
    // per SM: 12 warps x 32 lanes x 64 words x 4 B read and the same written per iteration
    const double bytes = 12.0 * 32 * 64 * 4 * iters;
    printf("{\"cuda_error\": \"%s\", \"mismatches\": %ld, \"cycles\": %lld, \"tmem_read_B_per_clk_per_SM\": %.1f, \"note\": \"read-modify-write, ld x16 + wait + st x16 + wait per chunk, 12 warps\"}\n",
           cudaGetErrorString(e), bad, c, bytes / (double)c);
    return 0;
}
