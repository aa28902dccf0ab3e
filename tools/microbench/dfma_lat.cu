// Micro-benchmark: DFMA issue rate on one SM sub-partition as a function of warps per scheduler and independent chains per warp
// (B200).  Not product code.  Prints cycles per DFMA per scheduler.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, double a, double b, int iters, long long* cyc) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 0.001 + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int i = 0; i < ILP; i++) x[i] = __fma_rn(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP>
void run(int warps_per_sched, double* out, long long* cyc) {
    const int iters = 2000;
    k<ILP><<<1, warps_per_sched * 4 * 32>>>(out, 1.0000001, 1e-9, iters, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / ((double)iters * 8 * ILP * warps_per_sched);
    printf("  {\"warps_per_scheduler\": %d, \"chains_per_warp\": %d, \"cycles_per_dfma_per_scheduler\": %.2f, \"cycles_per_dependent_step\": %.1f},\n",
           warps_per_sched, ILP, per, (double)c / ((double)iters * 8));
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    printf("[\n");
    for (int w = 1; w <= 4; w++) { run<1>(w, out, cyc); run<2>(w, out, cyc); run<3>(w, out, cyc); run<4>(w, out, cyc); run<6>(w, out, cyc); run<8>(w, out, cyc); }
    printf("  {\"cuda_error\": \"%s\"}\n]\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
