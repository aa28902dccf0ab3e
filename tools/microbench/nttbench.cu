// Micro-benchmark of the register-resident 32-point networks of ntt32.cuh under the occupancy of the product kernel
// (24 warps per SM, 80 registers): how close do the real code blocks get to the 15.6 butterflies/clk/SM that the
// FMA-heavy pipe allows (profiles/intpipe_r01.json)?  Not product code.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../rustfhe_b200/csrc/cmux_steps.cuh"
using namespace tfhe;

#define ITERS 400
// MODE 0: ct32 uniform twiddles; 1: ct32 row twiddles (smem); 2: gs32 row twiddles; 3: gs32 uniform;
// 4: full forward transform of a digit polynomial (p1a-like without rotate + p1b); 5: full inverse (rows, transpose, cols, lift)
template <int MODE, int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) k(uint32_t* out, uint32_t seed) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* twF = smem;
    uint32_t* twI = smem + 32 * TWB_STRIDE;
    uint32_t* tiles = smem + 2 * 32 * TWB_STRIDE;
    for (int t = threadIdx.x; t < 32 * TWB_STRIDE; t += blockDim.x) { twF[t] = g_fwdB[t]; twI[t] = g_invB[t]; }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t* S = tiles + warp * 1024;
    for (int t = lane; t < 1024; t += 32) S[t] = (t * 2654435761u + seed) % P;
    __syncthreads();
    uint32_t x[32];
#pragma unroll
    for (int r = 0; r < 32; r++) x[r] = (threadIdx.x * 2654435761u + r * 40503u + seed) % P;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) { ct32(x, TwUniform<false>());
#pragma unroll
            for (int r = 0; r < 32; r++) x[r] = csub(csub(x[r], 2u * P2), P2); }
        if (MODE == 1) { ct32(x, TwRow{twF + lane * TWB_STRIDE});
#pragma unroll
            for (int r = 0; r < 32; r++) x[r] = csub(csub(x[r], 2u * P2), P2); }
        if (MODE == 2) gs32(x, TwRow{twI + lane * TWB_STRIDE});
        if (MODE == 3) gs32(x, TwUniform<true>());
        if (MODE == 4) {
#pragma unroll
            for (int r = 0; r < 32; r++) x[r] = to_residue(gadget_digit(S[32 * r + lane] + x[r], 0x02084000u, 1));
            ct32(x, TwUniform<false>());
#pragma unroll
            for (int r = 0; r < 32; r++) S[swz(r, lane)] = x[r];
            __syncwarp();
            p1b(lane, S, twF);
            __syncwarp();
        }
        if (MODE == 5) {
#pragma unroll
            for (int q = 0; q < 8; q++) { const uint4 v = *reinterpret_cast<const uint4*>(S + swz_chunk(lane, q)); x[4*q] = csub(v.x + x[4*q], P2); x[4*q+1] = csub(v.y, P2); x[4*q+2] = csub(v.z, P2); x[4*q+3] = csub(v.w, P2); }
            gs32(x, TwRow{twI + lane * TWB_STRIDE});
#pragma unroll
            for (int q = 0; q < 8; q++) *reinterpret_cast<uint4*>(S + swz_chunk(lane, q)) = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
            __syncwarp();
            p2b(lane, S, 1, x);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 32; r++) S[32 * r + lane] = x[r] % P;
            __syncwarp();
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) r ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r ^ S[lane];
}

template <int MODE, int WARPS, int MINB>
void run(const char* name, int sms, uint32_t* out, double bfly_per_iter_per_thread) {
    const size_t smem = (size_t)(2 * 32 * TWB_STRIDE + WARPS * 1024) * 4;
    cudaFuncSetAttribute(k<MODE, WARPS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int blocks = sms * MINB * 4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, WARPS, MINB><<<blocks, WARPS * 32, smem>>>(out, 1u); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); k<MODE, WARPS, MINB><<<blocks, WARPS * 32, smem>>>(out, 2u + rep); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double bf = (double)blocks * WARPS * 32 * ITERS * bfly_per_iter_per_thread;
    const double per_clk_sm = bf / (best * 1e-3) / 1.965e9 / sms;
    printf("  \"%s_w%d_b%d\": {\"ms\": %.4f, \"butterflies_per_clk_per_sm\": %.2f, \"frac_of_15.6\": %.3f},\n", name, WARPS, MINB, best, per_clk_sm,
           per_clk_sm / 15.6);
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    uint32_t* out; cudaMalloc(&out, (size_t)sms * 16 * 1024 * 4);
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d,\n", p.name, sms);
    run<0, 24, 1>("ct32_uniform", sms, out, 80);
    run<1, 24, 1>("ct32_row", sms, out, 80);
    run<2, 24, 1>("gs32_row", sms, out, 80);
    run<3, 24, 1>("gs32_uniform", sms, out, 80);
    run<4, 24, 1>("forward_full", sms, out, 160);
    run<5, 24, 1>("inverse_full", sms, out, 160);
    run<0, 6, 3>("ct32_uniform", sms, out, 80);
    run<1, 6, 3>("ct32_row", sms, out, 80);
    run<2, 6, 3>("gs32_row", sms, out, 80);
    run<4, 6, 3>("forward_full", sms, out, 160);
    run<5, 6, 3>("inverse_full", sms, out, 160);
    run<0, 12, 2>("ct32_uniform", sms, out, 80);
    run<1, 12, 2>("ct32_row", sms, out, 80);
    run<2, 12, 2>("gs32_row", sms, out, 80);
    run<4, 12, 2>("forward_full", sms, out, 160);
    run<5, 12, 2>("inverse_full", sms, out, 160);
    cudaError_t err = cudaDeviceSynchronize();
    printf("  \"cuda_error\": \"%s\"\n}\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
