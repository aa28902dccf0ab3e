// Integer-pipe micro-benchmark for B200 (sm_100a): measures the sustained rate of the
// instructions the NTT butterflies are made of.  Output: one JSON object on stdout.
// Not product code; it provides P_int for the roofline.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
    uint32_t a[ILP], b = seed | 1u, c = seed * 3u + 7u;
    uint64_t w[ILP];
    double d[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 2654435761u + i + seed; w[i] = a[i]; d[i] = a[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MODE == 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)(w[i] >> 32)), "r"(b));
            if (MODE == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (MODE == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b), "r"(c));
            if (MODE == 5) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
            if (MODE == 6) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(1.0000001), "d"(0.5));
            if (MODE == 7) {  // Shoup lazy CT butterfly: x,y -> x+t, x-t+2p ; t = y*w - mulhi(y,w')*p
                uint32_t x = a[i], y = (uint32_t)w[i], q, t;
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(y), "r"(b));
                asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(y), "r"(c));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t) : "r"(q), "r"(0u - 536813569u));
                a[i] = x + t;
                w[i] = x - t + 2u * 536813569u;
            }
            if (MODE >= 9) {  // lazy butterfly + (MODE-8) extra independent ALU instructions: does the issue port keep up?
                uint32_t x = a[i], y = (uint32_t)w[i], q, t;
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(y), "r"(b));
                asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(y), "r"(c));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t) : "r"(q), "r"(0u - 536813569u));
                uint32_t e = (uint32_t)(w[i] >> 32);
                if (MODE >= 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e) : "r"(b), "r"(c));
                if (MODE >= 10) asm volatile("min.u32 %0, %0, %1;" : "+r"(e) : "r"(x));
                if (MODE >= 11) asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(e) : "r"(y), "r"(c));
                if (MODE >= 12) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(e) : "r"(x));
                a[i] = x + t;
                w[i] = ((uint64_t)e << 32) | (uint32_t)(x - t + 2u * 536813569u);
            }
            if (MODE == 8) {  // same + Harvey correction of x (min trick)
                uint32_t x = a[i], y = (uint32_t)w[i], q, t;
                x = min(x, x - 2u * 536813569u);
                asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(q) : "r"(y), "r"(b));
                asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(y), "r"(c));
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(t) : "r"(q), "r"(0u - 536813569u));
                a[i] = x + t;
                w[i] = x - t + 2u * 536813569u;
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r ^= a[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
double run(const char* name, int blocks, uint32_t* out, double opsPerIter) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 12345u); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, 12345u + rep); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double ops = (double)blocks * 256 * ITERS * ILP * opsPerIter;
    double rate = ops / (best * 1e-3);
    printf("  \"%s\": {\"ms\": %.4f, \"Gops_per_s\": %.1f},\n", name, best, rate / 1e9);
    return rate;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int blocks = sms * 8;
    uint32_t* out; cudaMalloc(&out, (size_t)blocks * 256 * 4);
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d,\n", p.name, sms, p.clockRate);
    run<0>("imad_lo", blocks, out, 1);
    run<1>("imad_hi", blocks, out, 1);
    run<2>("imad_wide_acc64", blocks, out, 1);
    run<3>("iadd", blocks, out, 1);
    run<4>("lop3", blocks, out, 1);
    run<5>("umin", blocks, out, 1);
    run<6>("dfma", blocks, out, 1);
    run<7>("shoup_ct_butterfly_lazy", blocks, out, 1);
    run<8>("shoup_ct_butterfly_harvey", blocks, out, 1);
    run<9>("butterfly_plus1_alu", blocks, out, 1);
    run<10>("butterfly_plus2_alu", blocks, out, 1);
    run<11>("butterfly_plus3_alu", blocks, out, 1);
    run<12>("butterfly_plus4_alu", blocks, out, 1);
    cudaError_t err = cudaDeviceSynchronize();
    printf("  \"cuda_error\": \"%s\"\n}\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
