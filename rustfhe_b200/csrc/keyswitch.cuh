// keyswitch.cuh -- TLWERep::identity_key_switch (tlwe.rs:43-73) as device kernels (included by engine.cu only).
#pragma once
#include "blind_rotate.cuh"   // smem_u32, LWE_N

// =====================================================================================================
// K6: key switch.  out[g] -= sum_{i,l : d != 0} KSK[i][l][d-1]  with d = 2-bit digit (i,l) of gate g.
// CTA = (tile of KS_GT gates) x (slice of 1024/KS_ISPLIT key indices); thread = one 16-byte column chunk of the
// 636-word rows (159 chunks).  Every KSK row is read once per CTA and applied to all gates of the tile; the digit
// is CTA-uniform so the select is a uniform branch.  Partial sums are merged with red.global.add.u32.
// =====================================================================================================
constexpr int KS_GT = 16;           // gates per thread group (accumulators live in registers: 16 x uint4)
#if !defined(KS_NG)
#define KS_NG 2
#endif
constexpr int KS_GROUPS = KS_NG;    // thread groups per CTA: they walk the same key rows, the second one hits L1
constexpr int KS_ISPLIT_MIN = 8;    // key indices are split over gridDim.y CTAs: 8 for large batches, up to 128 for small ones
constexpr int KS_ICHUNK = 1024 / KS_ISPLIT_MIN;   // largest slice of key indices one CTA walks
constexpr int KS_CHUNKS = (LWE_N + 1) / 4;  // 159
constexpr int KS_THREADS = 160;
__global__ void __launch_bounds__(KS_THREADS* KS_GROUPS) keyswitch_kernel(const uint4* __restrict__ ksk, const uint16_t* __restrict__ dig,
                                                                         uint32_t* __restrict__ out, long B, const int32_t* __restrict__ idxo) {
    __shared__ __align__(16) uint16_t dg[KS_GROUPS][KS_ICHUNK][KS_GT];
    const int grp = threadIdx.y;
    const long g0 = ((long)blockIdx.x * KS_GROUPS + grp) * KS_GT;
    const int ichunk = 1024 / (int)gridDim.y;
    const int i0 = blockIdx.y * ichunk;
    for (int t = threadIdx.x; t < ichunk * KS_GT; t += KS_THREADS) {
        const int g = t / ichunk, ii = t % ichunk;
        dg[grp][ii][g] = (g0 + g < B) ? dig[(size_t)(g0 + g) * 1024 + i0 + ii] : (uint16_t)0;
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= KS_CHUNKS || g0 >= B) return;
    uint4 acc[KS_GT];
#pragma unroll
    for (int g = 0; g < KS_GT; g++) acc[g] = make_uint4(0, 0, 0, 0);
    const uint4* base = ksk + (size_t)i0 * 8 * 3 * KS_CHUNKS + t;
#pragma unroll 1
    for (int ii = 0; ii < ichunk; ii++) {
        uint32_t dw[KS_GT / 2];
#pragma unroll
        for (int g = 0; g < KS_GT / 2; g++) dw[g] = reinterpret_cast<const uint32_t*>(dg[grp][ii])[g];  // two gates per word
#pragma unroll
        for (int l = 0; l < 8; l++) {
            const uint4* row = base + (size_t)(ii * 8 + l) * 3 * KS_CHUNKS;
            const uint4 r0 = __ldg(row), r1 = __ldg(row + KS_CHUNKS), r2 = __ldg(row + 2 * KS_CHUNKS);
#pragma unroll
            for (int g = 0; g < KS_GT; g++) {
                const uint32_t d = (dw[g >> 1] >> ((g & 1) * 16 + 14 - 2 * l)) & 3u;
                if (d == 1) { acc[g].x += r0.x; acc[g].y += r0.y; acc[g].z += r0.z; acc[g].w += r0.w; }
                else if (d == 2) { acc[g].x += r1.x; acc[g].y += r1.y; acc[g].z += r1.z; acc[g].w += r1.w; }
                else if (d == 3) { acc[g].x += r2.x; acc[g].y += r2.y; acc[g].z += r2.z; acc[g].w += r2.w; }
            }
        }
    }
#pragma unroll
    for (int g = 0; g < KS_GT; g++) {
        if (g0 + g >= B) break;
        uint32_t* o = out + (size_t)(idxo ? (long)idxo[g0 + g] : g0 + g) * (LWE_N + 1) + 4 * t;
        atomicAdd(o + 0, 0u - acc[g].x);
        atomicAdd(o + 1, 0u - acc[g].y);
        atomicAdd(o + 2, 0u - acc[g].z);
        atomicAdd(o + 3, 0u - acc[g].w);
    }
}
// ---- K6b: key switch with the key rows staged through shared memory, one WARP per gate ----
// The register-tile kernel above is ALU bound on its digit select (12 instructions per gate, digit and 16-byte chunk:
// the digit is CTA-uniform but every thread tests it).  Here a warp owns one gate and all 159 chunks of its output row
// (5 per lane): the digit picks the ROW ADDRESS in shared memory, so per (gate, digit) there are two instructions of
// select and five (LDS.128 + 4 adds) instead of 5 warps x 12.  The rows of a stage (4 levels x 3 multiples of one key
// index = 30 KB) are copied once per CTA by one bulk (TMA) copy into a 3-deep ring (one __syncthreads per stage) and consumed
// by the 16 gates of the tile.  Kept as TFHE_B200_KS_VARIANT=2; the shipped kernel is K6p below (the same rows behind a
// producer / consumer ring).
#if !defined(KS2_NG)
#define KS2_NG 16
#endif
constexpr int KS2_GATES = KS2_NG;                  // warps per CTA
constexpr int KS2_THREADS = KS2_GATES * 32;
#if !defined(KS2_LVDEF)
#define KS2_LVDEF 4
#endif
constexpr int KS2_LV = KS2_LVDEF;                  // levels per stage
constexpr int KS2_ROWS = KS2_LV * 3;               // rows per stage
constexpr int KS2_ROW_WORDS = LWE_N + 1;           // 636 words = 159 x 16 bytes: the rows of a stage are contiguous in the key, one bulk copy
constexpr int KS2_RING = 3;
constexpr int KS2_STAGE_WORDS = KS2_ROWS * KS2_ROW_WORDS;
constexpr size_t KS2_SMEM_BYTES = (size_t)KS2_RING * KS2_STAGE_WORDS * 4 + (size_t)KS_ICHUNK * KS2_GATES * 2 + KS2_RING * 8;
__global__ void __launch_bounds__(KS2_THREADS) keyswitch2_kernel(const uint4* __restrict__ ksk, const uint16_t* __restrict__ dig,
                                                                uint32_t* __restrict__ out, long B, const int32_t* __restrict__ idxo) {
    extern __shared__ __align__(16) uint32_t ks_smem[];
    uint32_t* ring = ks_smem;
    uint16_t* dg = reinterpret_cast<uint16_t*>(ks_smem + KS2_RING * KS2_STAGE_WORDS);   // [ichunk][16]
    uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(ks_smem) + KS2_SMEM_BYTES - KS2_RING * 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long g0 = (long)blockIdx.x * KS2_GATES;
    const int ichunk = 1024 / (int)gridDim.y;
    const int i0 = blockIdx.y * ichunk;
    const int nstages = ichunk * (8 / KS2_LV);
    for (int t = threadIdx.x; t < ichunk * KS2_GATES; t += KS2_THREADS) {
        const int g = t / ichunk, ii = t % ichunk;
        dg[ii * KS2_GATES + g] = (g0 + g < B) ? dig[(size_t)(g0 + g) * 1024 + i0 + ii] : (uint16_t)0;
    }
    // rows (key index i0 + k / SPI, KS2_LV levels, all three multiples) -> ring slot k % 3: ONE bulk (TMA) copy of 30 KB issued by
    // one thread, completing on the slot's mbarrier (the per-thread cp.async version spent 16 of its 58 instructions per gate,
    // index and level on the addresses of the 16-byte pieces)
    auto stage_in = [&](int k) {
        constexpr int SPI = 8 / KS2_LV;   // stages per key index
        const uint4* src = ksk + ((size_t)(i0 + k / SPI) * 8 + (size_t)(k % SPI) * KS2_LV) * 3 * KS_CHUNKS;
        bulk_fetch(ring + (k % KS2_RING) * KS2_STAGE_WORDS, src, KS2_STAGE_WORDS * 4, full + k % KS2_RING);
    };
    if (threadIdx.x == 0) {
        for (int r = 0; r < KS2_RING; r++) mbar_init(full + r, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        stage_in(0);
        if (nstages > 1) stage_in(1);
    }
    __syncthreads();   // digits staged, mbarriers initialised
    uint4 acc[5];
#pragma unroll
    for (int q = 0; q < 5; q++) acc[q] = make_uint4(0, 0, 0, 0);
    const bool live = g0 + warp < B;
#pragma unroll 1
    for (int k = 0; k < nstages; k++) {
        if (k > 0) __syncthreads();                               // everybody is done with stage k-1: its slot takes stage k+2
        if (threadIdx.x == 0 && k + 2 < nstages) stage_in(k + 2);
        mbar_wait(full + k % KS2_RING, (uint32_t)((k / KS2_RING) & 1));   // stage k has landed
        if (live) {
            constexpr int SPI = 8 / KS2_LV;
            const uint32_t d16 = dg[(k / SPI) * KS2_GATES + warp];
            const uint32_t* rows = ring + (k % KS2_RING) * KS2_STAGE_WORDS;
#pragma unroll
            for (int l = 0; l < KS2_LV; l++) {
                const uint32_t d = (d16 >> (14 - 2 * ((k % SPI) * KS2_LV + l))) & 3u;   // level 0 in bits 15:14
                if (d != 0) {
                    const uint4* r = reinterpret_cast<const uint4*>(rows + (l * 3 + (int)d - 1) * KS2_ROW_WORDS) + lane;
#pragma unroll
                    for (int q = 0; q < 5; q++) {
                        if (q < 4 || lane < KS_CHUNKS - 128) {
                            const uint4 v = r[32 * q];
                            acc[q].x += v.x; acc[q].y += v.y; acc[q].z += v.z; acc[q].w += v.w;
                        }
                    }
                }
            }
        }
    }
    if (!live) return;
    uint32_t* o = out + (size_t)(idxo ? (long)idxo[g0 + warp] : g0 + warp) * (LWE_N + 1);
#pragma unroll
    for (int q = 0; q < 5; q++) {
        const int c = lane + 32 * q;
        if (c < KS_CHUNKS) {
            atomicAdd(o + 4 * c + 0, 0u - acc[q].x);
            atomicAdd(o + 4 * c + 1, 0u - acc[q].y);
            atomicAdd(o + 4 * c + 2, 0u - acc[q].z);
            atomicAdd(o + 4 * c + 3, 0u - acc[q].w);
        }
    }
}
// ---- K6p: the same kernel as a producer / consumer pipeline ----
// K6b synchronises its 16 warps with one __syncthreads per stage (the slot of stage k - 1 takes stage k + 2) and every 16-gate
// tile stages its whole key slice from L2: 64 tiles x 62.5 MB = 4-6 GB per 1024 gates, 9.5 TB/s at 0.62 ms -- as close to the L2
// as to the shared-memory floor.  Here NG consumer warps (one gate each) and one producer warp share a ring of RING slots with
// full / empty mbarriers: a consumer waits for `full`, adds its rows, and its lane 0 arrives on `empty`; the producer waits for
// `empty` before it overwrites a slot.  No CTA-wide barrier in the loop: 0.54 ms per 1024 gates against 0.62.  More gates per CTA
// (24 or 31 with a ring of six, one CTA per SM: half the L2 traffic) measured the same 0.54-0.60 ms -- the kernel sits on its
// shared-memory floor (0.43 ms), not on the L2 (profiles/r02_keyswitch_pipeline_variants.log).
#if !defined(KSP_NG)
#define KSP_NG 16
#endif
#if !defined(KSP_NR)
#define KSP_NR 3
#endif
constexpr int KSP_GATES = KSP_NG, KSP_RING = KSP_NR;   // shipped shape: 16 gates + producer, ring of three: two CTAs per SM
template <int NG, int RING>
struct KsP {
    static constexpr int THREADS = (NG + 1) * 32;
    static constexpr size_t SMEM_BYTES = (size_t)RING * KS2_STAGE_WORDS * 4 + (size_t)KS_ICHUNK * NG * 2 + 2 * RING * 8;
};
template <int NG, int RING>
__global__ void __launch_bounds__((NG + 1) * 32) keyswitch_p_kernel(const uint4* __restrict__ ksk, const uint16_t* __restrict__ dig,
                                                                    uint32_t* __restrict__ out, long B, const int32_t* __restrict__ idxo) {
    extern __shared__ __align__(16) uint32_t ks_smem[];
    uint32_t* ring = ks_smem;
    uint16_t* dg = reinterpret_cast<uint16_t*>(ks_smem + RING * KS2_STAGE_WORDS);   // [ichunk][NG]
    uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(ks_smem) + KsP<NG, RING>::SMEM_BYTES - 2 * RING * 8);
    uint64_t* empty = full + RING;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long g0 = (long)blockIdx.x * NG;
    // the 1024 key indices are cut into gridDim.y nearly equal slices (any count from KS_ISPLIT_MIN to 128, not only powers of two:
    // the host picks the count that fills the SMs in whole waves)
    const int i0 = (int)(((long)blockIdx.y * 1024) / (long)gridDim.y);
    const int ichunk = (int)((((long)blockIdx.y + 1) * 1024) / (long)gridDim.y) - i0;
    const int nstages = ichunk * (8 / KS2_LV);
    constexpr int SPI = 8 / KS2_LV;   // stages per key index
    for (int t = threadIdx.x; t < ichunk * NG; t += (NG + 1) * 32) {
        const int g = t / ichunk, ii = t % ichunk;
        dg[ii * NG + g] = (g0 + g < B) ? dig[(size_t)(g0 + g) * 1024 + i0 + ii] : (uint16_t)0;
    }
    if (threadIdx.x == 0) {
        for (int r = 0; r < RING; r++) { mbar_init(full + r, 1); mbar_init(empty + r, NG); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();   // digits staged, mbarriers initialised
    if (warp == NG) {   // producer
        if (lane == 0) {
#pragma unroll 1
            for (int k = 0; k < nstages; k++) {
                const int s = k % RING;
                if (k >= RING) mbar_wait(empty + s, (uint32_t)((k / RING - 1) & 1));   // every consumer has left the slot's previous stage
                const uint4* src = ksk + ((size_t)(i0 + k / SPI) * 8 + (size_t)(k % SPI) * KS2_LV) * 3 * KS_CHUNKS;
                bulk_fetch(ring + s * KS2_STAGE_WORDS, src, KS2_STAGE_WORDS * 4, full + s);
            }
        }
        return;
    }
    uint4 acc[5];
#pragma unroll
    for (int q = 0; q < 5; q++) acc[q] = make_uint4(0, 0, 0, 0);
    const bool live = g0 + warp < B;   // warps without a gate still hand their slots back
#pragma unroll 1
    for (int k = 0; k < nstages; k++) {
        const int s = k % RING;
        mbar_wait(full + s, (uint32_t)((k / RING) & 1));
        if (live) {
            const uint32_t d16 = dg[(k / SPI) * NG + warp];
            const uint32_t* rows = ring + s * KS2_STAGE_WORDS;
#pragma unroll
            for (int l = 0; l < KS2_LV; l++) {
                const uint32_t d = (d16 >> (14 - 2 * ((k % SPI) * KS2_LV + l))) & 3u;   // level 0 in bits 15:14
                if (d != 0) {
                    const uint4* r = reinterpret_cast<const uint4*>(rows + (l * 3 + (int)d - 1) * KS2_ROW_WORDS) + lane;
#pragma unroll
                    for (int q = 0; q < 5; q++) {
                        if (q < 4 || lane < KS_CHUNKS - 128) {
                            const uint4 v = r[32 * q];
                            acc[q].x += v.x; acc[q].y += v.y; acc[q].z += v.z; acc[q].w += v.w;
                        }
                    }
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + s);
    }
    if (!live) return;
    uint32_t* o = out + (size_t)(idxo ? (long)idxo[g0 + warp] : g0 + warp) * (LWE_N + 1);
#pragma unroll
    for (int q = 0; q < 5; q++) {
        const int c = lane + 32 * q;
        if (c < KS_CHUNKS) {
            atomicAdd(o + 4 * c + 0, 0u - acc[q].x);
            atomicAdd(o + 4 * c + 1, 0u - acc[q].y);
            atomicAdd(o + 4 * c + 2, 0u - acc[q].z);
            atomicAdd(o + 4 * c + 3, 0u - acc[q].w);
        }
    }
}
// prepares the key-switch inputs from explicit level-1 samples (step-level entry tfhe_b200_keyswitch_batch)
__global__ void lwe1_prepare_kernel(const uint32_t* __restrict__ lwe1, uint16_t* __restrict__ dig, uint32_t* __restrict__ out, long B) {
    const long g = blockIdx.x;
    if (g >= B) return;
    const uint32_t* src = lwe1 + (size_t)g * 1025;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) dig[(size_t)g * 1024 + i] = (uint16_t)((src[1 + i] + 0x8000u) >> 16);
    for (int c = threadIdx.x; c <= LWE_N; c += blockDim.x) out[(size_t)g * (LWE_N + 1) + c] = (c == 0) ? src[0] : 0u;
}

