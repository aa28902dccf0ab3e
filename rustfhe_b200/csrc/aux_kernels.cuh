// aux_kernels.cuh -- exact negacyclic product micro-entry (math.rs:337-347) and the device-side key generation / encryption /
// decryption / sample-extract kernels (included by engine.cu only).
#pragma once
#include "blind_rotate.cuh"
#include "tfhe_rng.cuh"

// =====================================================================================================
// exact negacyclic product a (torus) * d (small ints): one warp per product, 7 transforms
// =====================================================================================================
constexpr int PM_WARPS = 2;
// product g reads a = A + g*a_stride, d = D + g*d_stride (d_stride 0: one multiplier shared by the batch) and writes
// out + g*o_stride (accumulate: += instead of =)
__global__ void __launch_bounds__(PM_WARPS * 32) polymul_kernel(const uint32_t* __restrict__ A, const int32_t* __restrict__ D,
                                                               uint32_t* __restrict__ out, long B, long a_stride, long d_stride,
                                                               long o_stride, int accumulate) {
    __shared__ __align__(16) uint32_t twF[32 * TWB_STRIDE];
    __shared__ __align__(16) uint32_t twI[32 * TWB_STRIDE];
    __shared__ __align__(16) uint32_t scratch[PM_WARPS][2][TILE_WORDS];
    for (int t = threadIdx.x; t < 32 * TWB_STRIDE; t += blockDim.x) { twF[t] = g_fwdB[t]; twI[t] = g_invB[t]; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long g = (long)blockIdx.x * PM_WARPS + warp;
    if (g >= B) return;
    const uint32_t* a = A + (size_t)g * a_stride;
    const int32_t* d = D + (size_t)g * d_stride;
    uint32_t* S = scratch[warp][0];
    uint32_t* T = scratch[warp][1];
    uint32_t dh[32], x[32], res[32];
#pragma unroll
    for (int r = 0; r < 32; r++) x[r] = to_residue(d[32 * r + lane]);
    fwd_cols(lane, x, S);
    __syncwarp();
    fwd_rows(lane, S, twF, dh);  // spectrum of d, row layout, in [0,p)
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 32; r++) res[r] = 0;
    for (int part = 0; part < 3; part++) {
        key_cols(lane, a, part, S);
        __syncwarp();
        key_rows(lane, S, twF, T);  // [q][lane][4], scaled by 2^32/N
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const uint4 b = *reinterpret_cast<const uint4*>(T + (q * 32 + lane) * 4);
            x[4 * q] = redc64((uint64_t)dh[4 * q] * b.x);
            x[4 * q + 1] = redc64((uint64_t)dh[4 * q + 1] * b.y);
            x[4 * q + 2] = redc64((uint64_t)dh[4 * q + 2] * b.z);
            x[4 * q + 3] = redc64((uint64_t)dh[4 * q + 3] * b.w);
        }
        inv_rows(lane, x, twI, S);
        __syncwarp();
        p2b(lane, S, part, x);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 32; r++) res[r] += x[r];
    }
    uint32_t* o = out + (size_t)g * o_stride;
#pragma unroll
    for (int r = 0; r < 32; r++) o[32 * r + lane] = accumulate ? o[32 * r + lane] + res[r] : res[r];
}

// =====================================================================================================
// Device-side key generation and encryption.  Same seeded counter generator and the same operation
// order as the host keygen (hostkeys.cpp, tfhe_rng.cuh): the device keys are bit-identical to the host keys.
// Reference: BootstrappingKey::new (tfhe.rs:119-126), TRGSW/TRLWE encrypt (trgsw.rs:117-139,213-229; trlwe.rs:127-137),
// KeySwitchingKey::new (tlwe.rs:247-277), TLWE encrypt / decrypt (tlwe.rs:213-240).
// =====================================================================================================
using tfhe_rng::Rng;
// rows of the bootstrapping key before the a*s product: A = uniform, B = noise      bk: [n][2l][2][N], poly 0 = B, poly 1 = A
__global__ void bk_fill_kernel(uint32_t* __restrict__ bk, const tfhe_rng::RngKey seed, long nwords /* = rows * N */) {
    const Rng ra(seed, tfhe_rng::BK_A), re(seed, tfhe_rng::BK_E);
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < nwords; t += (long)gridDim.x * blockDim.x) {
        const long row = t >> 10;
        const int k = (int)(t & 1023);
        bk[(size_t)(row * 2 + 1) * 1024 + k] = ra.u32((uint64_t)t);
        bk[(size_t)(row * 2 + 0) * 1024 + k] = re.gauss((uint64_t)t, tfhe_rng::SCALE_BK);
    }
}
// gadget term of TRGSW_{s1}(s0_i): s0_i / Bg^(j+1) on B[0] of rows j < l and on A[0] of rows l + j (trgsw.rs:213-229)
__global__ void bk_gadget_kernel(uint32_t* __restrict__ bk, const uint8_t* __restrict__ s0, int rows) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= rows) return;
    const int i = row / 6, j = row % 6;
    const uint32_t mu = (uint32_t)s0[i] << (32 - 6 * ((j % 3) + 1));
    bk[(size_t)(row * 2 + (j < 3 ? 0 : 1)) * 1024] += mu;
}
// one warp per LWE row under s0: a = uniform, b = <a, s0> + noise + message.
//   mode 0: key-switching key, row id = (i, l, d-1), message = d * s1_i / 2^(2(l+1))      (tlwe.rs:247-283)
//   mode 1: encryption of bits[g], row id = ct_index0 + g, message = +-1/8                  (tlwe.rs:181-186,213-228)
__global__ void lwe_rows_kernel(uint32_t* __restrict__ out, long rows, const tfhe_rng::RngKey seed, uint64_t index0, const uint8_t* __restrict__ s0,
                                const uint8_t* __restrict__ s1, const uint8_t* __restrict__ bits, int mode) {
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const Rng ra(seed, mode == 0 ? tfhe_rng::KSK_A : tfhe_rng::ENC_A), re(seed, mode == 0 ? tfhe_rng::KSK_E : tfhe_rng::ENC_E);
    const uint64_t id = index0 + (uint64_t)row;
    uint32_t* ct = out + (size_t)row * (LWE_N + 1);
    uint32_t part = 0;
    for (int c = lane; c < LWE_N; c += 32) {
        const uint32_t av = ra.u32(id * LWE_N + c);
        ct[1 + c] = av;
        if (s0[c]) part += av;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
    if (lane == 0) {
        uint32_t msg;
        if (mode == 0) {
            const int i = (int)(row / 24), l = (int)((row / 3) % 8), d = (int)(row % 3) + 1;
            msg = (uint32_t)(d * s1[i]) << (32 - 2 * (l + 1));
        } else {
            msg = bits[row] ? 0x20000000u : 0xE0000000u;
        }
        ct[0] = msg + re.gauss(id, tfhe_rng::SCALE_LV0) + part;
    }
}
// phase = b - <a, s0> and the decoded bit (tlwe.rs:187-194,230-240); one warp per ciphertext
__global__ void lwe_phase_kernel(const uint32_t* __restrict__ ct, long rows, const uint8_t* __restrict__ s0, uint32_t* __restrict__ phase,
                                 uint8_t* __restrict__ bits) {
    const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const uint32_t* c = ct + (size_t)row * (LWE_N + 1);
    uint32_t part = 0;
    for (int k = lane; k < LWE_N; k += 32) if (s0[k]) part += c[1 + k];
#pragma unroll
    for (int o = 16; o; o >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, o);
    if (lane == 0) {
        const uint32_t ph = c[0] - part;
        if (phase) phase[row] = ph;
        if (bits) bits[row] = ((float)ph * (1.0f / 4294967296.0f)) < 0.5f ? 1 : 0;   // torus2binary, math.rs:684-690
    }
}
// TRLWERep::sample_extract_index(index) (trlwe.rs:110-121): b' = b[index]; a'_i = a[index-i] (i <= index), -a[N+index-i] otherwise
__global__ void sample_extract_kernel(const uint32_t* __restrict__ trlwe, uint32_t* __restrict__ out, long B, int index) {
    const long g = blockIdx.x;
    if (g >= B) return;
    const uint32_t* b = trlwe + (size_t)g * 2048;
    const uint32_t* a = b + 1024;
    uint32_t* o = out + (size_t)g * 1025;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) o[1 + i] = (i <= index) ? a[index - i] : 0u - a[1024 + index - i];
    if (threadIdx.x == 0) o[0] = b[index];
}

// wires[out[g]] = rows[g]: one level's ciphertexts, gathered from every device of a group, into the wire table (one warp per row)
__global__ void wire_scatter_kernel(const uint4* __restrict__ rows, const int32_t* __restrict__ idxo, uint32_t* __restrict__ wires, long n) {
    const long g = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    const uint4* src = rows + (size_t)g * ((LWE_N + 1) / 4);
    uint4* dst = reinterpret_cast<uint4*>(wires + (size_t)idxo[g] * (LWE_N + 1));
    for (int c = threadIdx.x & 31; c < (LWE_N + 1) / 4; c += 32) dst[c] = src[c];
}

