// Issue-port micro-benchmark (B200): cycles per warp for fixed instruction mixes, 64 warps/SM, ILP 8.  Not product code.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define ITERS 4096
#define ILP 8
// each MODE: NI = instructions per chain-iteration
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t seed) {
    uint32_t a[ILP], e[ILP], f[ILP], b = seed | 1u, c = seed * 3u + 7u;
#pragma unroll
    for (int i = 0; i < ILP; i++) { a[i] = threadIdx.x * 2654435761u + i + seed; e[i] = a[i] ^ 0x55u; f[i] = a[i] + 77u; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
#define IMAD asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b), "r"(c));
#define IMADHI asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b));
#define IADD3E asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(e[i]) : "r"(b), "r"(c));
#define IADD3F asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(f[i]) : "r"(c), "r"(b));
#define LOP3E asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(e[i]) : "r"(b), "r"(c));
#define MINF asm volatile("min.u32 %0, %0, %1;" : "+r"(f[i]) : "r"(c));
            if (MODE == 0) { IMAD }
            if (MODE == 1) { IMAD IADD3E }
            if (MODE == 2) { IMAD IADD3E IADD3F }
            if (MODE == 3) { IMAD IADD3E IADD3F IADD3E }
            if (MODE == 4) { IMADHI IADD3E }
            if (MODE == 5) { IMADHI IADD3E IADD3F }
            if (MODE == 6) { IMADHI IADD3E IADD3F IADD3E }
            if (MODE == 7) { IMADHI IADD3E IADD3F IADD3E IADD3F }
            if (MODE == 8) { IMAD LOP3E }
            if (MODE == 9) { IMAD LOP3E MINF }
            if (MODE == 10) { IADD3E }
            if (MODE == 11) { IADD3E LOP3E }
            if (MODE == 12) { IMAD MINF }
            if (MODE == 13) { IMAD MINF IADD3E }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) r ^= a[i] ^ e[i] ^ f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
template <int MODE>
void run(const char* name, int blocks, uint32_t* out, int ni, int sms) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(out, 12345u); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, 12345u + rep); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // warp-iterations per SMSP = blocks*8 warps / (sms*4) * ITERS*ILP
    const double wi = (double)blocks * 8 / (sms * 4.0) * ITERS * ILP;
    const double cyc = best * 1e-3 * 1.965e9 / wi;
    printf("  \"%s\": {\"instr\": %d, \"cycles_per_warp_iter\": %.3f, \"ipc_smsp\": %.3f},\n", name, ni, cyc, ni / cyc);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int blocks = sms * 8;
    uint32_t* out; cudaMalloc(&out, (size_t)blocks * 256 * 4);
    printf("{\n");
    run<0>("imad", blocks, out, 1, sms);
    run<1>("imad+iadd3", blocks, out, 2, sms);
    run<2>("imad+2iadd3", blocks, out, 3, sms);
    run<3>("imad+3iadd3", blocks, out, 4, sms);
    run<4>("imadhi+iadd3", blocks, out, 2, sms);
    run<5>("imadhi+2iadd3", blocks, out, 3, sms);
    run<6>("imadhi+3iadd3", blocks, out, 4, sms);
    run<7>("imadhi+4iadd3", blocks, out, 5, sms);
    run<8>("imad+lop3", blocks, out, 2, sms);
    run<9>("imad+lop3+min", blocks, out, 3, sms);
    run<10>("iadd3", blocks, out, 1, sms);
    run<11>("iadd3+lop3", blocks, out, 2, sms);
    run<12>("imad+min", blocks, out, 2, sms);
    run<13>("imad+min+iadd3", blocks, out, 3, sms);
    printf("  \"cuda_error\": \"%s\"\n}\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
