// Micro-benchmark: Shoup butterfly with the quotient from IMAD.HI vs from one DFMA on the FP64 pipe (B200).  Not product code.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
#define ILP 8
constexpr uint32_t P = 536856577u;
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed, double winv, double cmagic) {
    uint32_t a[ILP], y[ILP], w = seed | 1u, ws = seed * 3u + 7u;
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 2654435761u + i + seed; y[i] = a[i] ^ 0x5555u; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            uint32_t q, t;
            if (MODE == 0) {
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(y[i]), "r"(ws));
            } else if (MODE == 1) {   // q = low word of fma((2^52 + y), winv, C)
                const double yd = __hiloint2double(0x43300000, (int)y[i]);
                const double r = __fma_rn(yd, winv, cmagic);
                q = (uint32_t)__double2loint(r);
            } else {                  // explicit conversion: (2^52+y) - 2^52, then fma with magic
                const double yd = __dadd_rn(__hiloint2double(0x43300000, (int)y[i]), -4503599627370496.0);
                const double r = __fma_rn(yd, winv, 6755399441055744.0);
                q = (uint32_t)__double2loint(r);
            }
            asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(y[i]), "r"(w));
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t) : "r"(q), "r"(0u - P));
            const uint32_t x = a[i];
            a[i] = x + t;
            y[i] = x - t + 2u * P;
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r ^= a[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name, int blocks, uint32_t* out, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 12345u, 0.37, 6755399441055744.0); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, 12345u + rep, 0.37, 6755399441055744.0); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double wi = (double)blocks * 8 / (sms * 4.0) * ITERS * ILP;
    printf("  \"%s\": {\"cycles_per_warp_butterfly\": %.3f},\n", name, best * 1e-3 * 1.965e9 / wi);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int blocks = sms * 8;
    uint32_t* out; cudaMalloc(&out, (size_t)blocks * 256 * 4);
    printf("{\n");
    run<0>("shoup_imad_hi", blocks, out, sms);
    run<1>("shoup_dfma_fused_magic", blocks, out, sms);
    run<2>("shoup_dadd_dfma", blocks, out, sms);
    printf("  \"cuda_error\": \"%s\"\n}\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
