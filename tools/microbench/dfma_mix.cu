// Micro-benchmark: does a DFMA leave an issue slot free?  Per iteration every warp issues ND independent DFMAs and NI independent
// integer instructions (IADD3 / LOP3 on the ALU pipe) or NF FFMAs; cycles per iteration per scheduler.  Not product code.
#include <cstdio>
#include <cuda_runtime.h>
template <int ND, int NI, int NF>
__global__ void k(double* out, double a, double b, int iters, long long* cyc) {
    double x[8]; unsigned v[8]; float f[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 0.001 + i; v[i] = threadIdx.x + i; f[i] = threadIdx.x * 0.5f + i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (i < ND) x[i] = __fma_rn(x[i], a, b);
                if (i < NI) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(v[(i + 1) & 7] | 1u), "r"(0x5555u + it));
                if (i < NF) f[i] = __fmaf_rn(f[i], 1.0001f, 0.5f);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i] + v[i] + f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ND, int NI, int NF>
void run(int wps, double* out, long long* cyc) {
    const int iters = 2000;
    k<ND, NI, NF><<<1, wps * 4 * 32>>>(out, 1.0000001, 1e-9, iters, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("  {\"warps_per_scheduler\": %d, \"dfma\": %d, \"alu\": %d, \"ffma\": %d, \"cycles_per_group_per_scheduler\": %.2f, \"issue_slots_per_group\": %d},\n", wps, ND, NI,
           NF, (double)c / (iters * 4.0), wps * (ND + NI + NF));
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    printf("[\n");
    for (int w = 1; w <= 3; w++) {
        run<4, 0, 0>(w, out, cyc); run<4, 4, 0>(w, out, cyc); run<4, 8, 0>(w, out, cyc); run<4, 0, 4>(w, out, cyc); run<4, 4, 4>(w, out, cyc);
        run<0, 8, 0>(w, out, cyc); run<0, 0, 8>(w, out, cyc); run<8, 0, 0>(w, out, cyc); run<8, 8, 0>(w, out, cyc);
    }
    printf("  {\"cuda_error\": \"%s\"}\n]\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
