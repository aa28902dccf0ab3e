// blind_rotate_f64l2.cuh -- K5FL2, the LATENCY shape of the FFT64 mode (included by engine.cu only).
//   gate pre-combination + 635 x CMUX + sample extract (tfhe.rs:27-113, trgsw.rs:264-322, trlwe.rs:110-121)
// One gate per CTA (one SM), twelve warps.  The critical path of a CMUX is  source words -> one forward transform -> products ->
// sum -> one inverse transform -> accumulate;  with one warp per transform (K5FL) a transform takes about 3 000 cycles, almost
// all of it dependency latency.  Here a transform runs on TWO warps, 8 complex values per thread, as three radix-8 passes
// (three butterfly stages each, fully in registers) with two transposes through shared memory between them (64-thread named
// barriers): 216 FP64 operations per thread and transform instead of 432 + the lane-pair exchange.
//   forward (digit (pw, dw) on warp pair 3 pw + dw), thread t < 64:
//     pass 1 : j = t + 64 e, e < 8      stages 0..2 on e   (warp-uniform twiddles, the constant bank)
//     pass 2 : j = 64 hi + 8 m + lo     stages 3..5 on m   (thread = 8 hi + lo; twiddles per hi: 4 loads from a 512-byte table)
//     pass 3 : p = 8 tt + e''           stages 6..8 on e'' (thread tt; 4 loads from a 4 KB table)
//   The spectrum positions are those of the one-warp transform (in-place Cooley-Tukey: position p holds psi^(1 + 4 bitrev9 p)), so
//   the same transformed key values are used, re-laid-out per polynomial as [register 8][thread 64] (f64l2_key_layout_kernel): in
//   the one-warp layout a quarter warp would read two rows 4 KB apart, a two-way bank conflict on every key load.
//   The 96 KB of key of a step arrive by ONE bulk (TMA) copy, requested as soon as the previous step's products are done, i.e.
//   a whole inverse transform ahead (loads into registers were tried first: ptxas sinks them next to their use and the products
//   then wait 3 000 cycles for L2).  Each pair multiplies its spectrum by its two key polynomials and leaves the products in its
//   own transpose buffers; after a CTA barrier pairs 0 and 1 add the six products of output polynomial 0 / 1, meet at a
//   128-thread barrier (the buffers are free now) and run the inverse (decimation in time: pass 1' on p[2:0], pass 2' on p[5:3],
//   pass 3' on p[8:6], untwist, exact rounding).
// Results are bit-identical to the other kernels' (every product rounds to the exact integer).
#pragma once
#include "blind_rotate_f64.cuh"

#if !defined(F64L2_SHARED_U)
#define F64L2_SHARED_U 0   // 1: the masked source words are computed once per step by the whole CTA (one more CTA barrier): measured 2 % slower
#endif
#if !defined(F64L2_KEYREG)
#define F64L2_KEYREG 1     // 1: a thread's 16 key values are read from the key buffer before pass 3 (latency under the pass)
#endif
constexpr int F64L2_PAIRS = 6;
constexpr int F64L2_THREADS = F64L2_PAIRS * 64;
static __device__ const double g_l2_fwd2[4 * 8 * 2] = {FFT64_L2_FWD2_LIST};
static __device__ const double g_l2_fwd3[4 * 64 * 2] = {FFT64_L2_FWD3_LIST};
static __device__ const double g_l2_inv2[4 * 8 * 2] = {FFT64_L2_INV2_LIST};
static __device__ const double g_l2_inv3[4 * 64 * 2] = {FFT64_L2_INV3_LIST};
static __device__ const double g_l2_untw[8 * 64 * 2] = {FFT64_L2_UNTWIST_LIST};
constexpr int F64L2_TAB_ELEMS = 32 + 256 + 32 + 256 + 512;   // cd16 elements, in the order above
constexpr int F64L2_KEY_BYTES = (int)(F64_STEP_ELEMS * sizeof(cd16));   // 98 304: the key of one step
constexpr int F64L2_SMEM_BYTES = F64L2_TAB_ELEMS * 16 + 2 * 1024 * 4 /*acc*/ + F64L2_PAIRS * 2 * 512 * 16 /*transpose buffers A, B per pair = products*/ +
                                 F64L2_KEY_BYTES + 2 * 1024 * 4 /*masked source words*/ + 640 * 2 /*abar*/ + 16 /*mbarrier*/;
static_assert(F64L2_SMEM_BYTES <= 227 * 1024, "one gate must fit the shared memory of one SM");

// accumulator word += v as one shared-memory reduction: nothing to wait for (a load + add + store has the load's latency in the step's chain)
__device__ __forceinline__ void l2_red_add(uint32_t* p, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory"); }
__device__ __forceinline__ void l2_store(cd16* p, const cd& v) { cd16 t; t.re = v.re; t.im = v.im; *p = t; }
__device__ __forceinline__ void l2_load(const cd16* p, cd& v) { const cd16 t = *p; v.re = t.re; v.im = t.im; }

__global__ void __launch_bounds__(F64L2_THREADS, 1) blind_rotate_f64_latency2_kernel(const BrArgs a, const cd16* __restrict__ key) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd16* tab = reinterpret_cast<cd16*>(smem_raw);
    const cd16* tf2 = tab;            // [4][8]
    const cd16* tf3 = tab + 32;       // [4][64]
    const cd16* ti2 = tab + 288;      // [4][8]
    const cd16* ti3 = tab + 320;      // [4][64]
    const cd16* tut = tab + 576;      // [8][64]
    uint32_t* acc = reinterpret_cast<uint32_t*>(tab + F64L2_TAB_ELEMS);
    cd16* scratch = reinterpret_cast<cd16*>(acc + 2048);
    cd16* prod = scratch;                            // [row j][output o][register][thread]: pair j's products replace its transpose buffers
    cd16* keybuf = scratch + F64L2_PAIRS * 2 * 512;  // [row j][output o][register 16][lane 32]: the key of the current step
    uint32_t* uw = reinterpret_cast<uint32_t*>(keybuf + F64_STEP_ELEMS);   // ((X^abar acc - acc) + mask) ^ mask of both polynomials, once per step
    uint16_t* abar = reinterpret_cast<uint16_t*>(uw + 2048);
    uint64_t* kfull = reinterpret_cast<uint64_t*>(abar + 640);
    const int pair = threadIdx.x >> 6, t = threadIdx.x & 63;
    const int pw = pair / 3, dw = pair - 3 * pw;
    cd16* bufA = scratch + (size_t)pair * 1024;
    cd16* bufB = bufA + 512;
    const int bar_id = 1 + pair;
    const long gate = blockIdx.x;
    const int nsteps = a.nsteps;
    {
        double* d = reinterpret_cast<double*>(tab);
        for (int k = threadIdx.x; k < 64; k += blockDim.x) { d[k] = g_l2_fwd2[k]; d[2 * 288 + k] = g_l2_inv2[k]; }
        for (int k = threadIdx.x; k < 512; k += blockDim.x) { d[2 * 32 + k] = g_l2_fwd3[k]; d[2 * 320 + k] = g_l2_inv3[k]; }
        for (int k = threadIdx.x; k < 1024; k += blockDim.x) d[2 * 576 + k] = g_l2_untw[k];
    }
    // ---- prologue: gate pre-combination (tfhe.rs:27-71), rounding of (b, a) (tfhe.rs:97,107-108), acc_0 ----
    {
        uint32_t* lin = reinterpret_cast<uint32_t*>(scratch);
        const bool second = gate >= a.split;
        const long gsrc = second ? gate - a.split : gate;
        const uint32_t* q0 = second ? a.in0b : a.in0;
        const uint32_t* q1 = second ? a.in1b : a.in1;
        uint32_t k0 = (uint32_t)(second ? a.c0b : a.c0), k1 = (uint32_t)(second ? a.c1b : a.c1), kb = second ? a.cbb : a.cb;
        if (a.ops) gate_coeffs(a.ops[gate], a.mu, k0, k1, kb);
        const uint32_t* p0 = q0 + (size_t)(a.idx0 ? (long)a.idx0[gate] : gsrc) * (LWE_N + 1);
        const uint32_t* p1 = (q1 && k1 != 0) ? q1 + (size_t)(a.idx1 ? (long)a.idx1[gate] : gsrc) * (LWE_N + 1) : nullptr;
        for (int c = threadIdx.x; c <= LWE_N; c += F64L2_THREADS) {
            uint32_t v = k0 * p0[c];
            if (p1) v += k1 * p1[c];
            if (c == 0) v += kb;
            lin[c] = v;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 640; i += F64L2_THREADS) abar[i] = i < LWE_N ? (uint16_t)((lin[1 + i] + (1u << 20)) >> 21) : (uint16_t)0;   // round
        const uint32_t bbar = lin[0] >> 21;                                                                                                      // floor
        const uint32_t nrot = (2048u - bbar) & 2047u;
        for (int k = threadIdx.x; k < 1024; k += F64L2_THREADS) {
            const bool neg = ((uint32_t)k < (nrot & 1023u)) != (nrot >= 1024u);
            acc[k] = neg ? 0u - a.mu : a.mu;
            acc[1024 + k] = 0;
        }
    }
    if (threadIdx.x == 0) {
        mbar_init(kfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (nsteps > 0) bulk_fetch(keybuf, key, F64L2_KEY_BYTES, kfull);
    }
    __syncthreads();

    // per-thread twiddles: constant for the whole kernel (16 + 16 registers... 8 values x 4 words each way)
    const int hi3 = t >> 3, lo3 = t & 7;
    cd16 wf2[4], wf3[4];   // forward twiddles: resident; the inverse ones are loaded by the two pairs that need them
#pragma unroll
    for (int k = 0; k < 4; k++) { wf2[k] = tf2[k * 8 + hi3]; wf3[k] = tf3[k * 64 + t]; }
    // this thread's 8 spectrum points p = 8 t + e in the [register 16][lane 32] key layout: (p & 15) * 32 + (p >> 4)
    const cd16* kp = keybuf + (size_t)(2 * pair) * F64_CHUNK_ELEMS + t;   // [register e][thread t]
    const uint32_t* A = (F64L2_SHARED_U ? uw : acc) + pw * 1024;
    const int sh = 6 * dw;

#pragma unroll 1
    for (int i = 0; i < nsteps; i++) {
        cd x[8];
#if F64L2_SHARED_U
        {   // the masked source words of both polynomials, computed ONCE by the whole CTA (the three pairs of a polynomial need the same)
            const uint32_t ab = abar[i];
            for (int k = threadIdx.x; k < 2048; k += F64L2_THREADS)
                uw[k] = add_alu(rot_diff(acc + (k & 1024), (uint32_t)(k & 1023), ab), a.mask) ^ a.mask;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 8; e++) {   // folded input z_j = digit(u_j) + i digit(u_{j+512}), j = t + 64 e, digits times 4 (key scale)
            x[e].re = (double)((((int32_t)(A[t + 64 * e] << sh)) >> 24) & ~3);
            x[e].im = (double)((((int32_t)(A[512 + t + 64 * e] << sh)) >> 24) & ~3);
        }
#else
        {
            const uint32_t ab = abar[i];
#pragma unroll
            for (int e = 0; e < 8; e++) {   // folded input z_j = digit(u_j) + i digit(u_{j+512}), j = t + 64 e, digits times 4 (key scale)
                const uint32_t ur = add_alu(rot_diff(A, (uint32_t)(t + 64 * e), ab), a.mask) ^ a.mask;
                const uint32_t ui = add_alu(rot_diff(A, (uint32_t)(512 + t + 64 * e), ab), a.mask) ^ a.mask;
                x[e].re = (double)((((int32_t)(ur << sh)) >> 24) & ~3);
                x[e].im = (double)((((int32_t)(ui << sh)) >> 24) & ~3);
            }
        }
#endif
        l2_fwd_pass1(x);
#pragma unroll
        for (int e = 0; e < 8; e++) l2_store(bufA + t + 64 * e, x[e]);
        bar_sync(bar_id, 64);
#pragma unroll
        for (int m = 0; m < 8; m++) l2_load(bufA + 64 * hi3 + 8 * m + lo3, x[m]);
        l2_fwd_pass23(x, wf2);
#pragma unroll
        for (int m = 0; m < 8; m++) l2_store(bufB + 64 * hi3 + 8 * m + (lo3 ^ m), x[m]);
        bar_sync(bar_id, 64);
#pragma unroll
        for (int e = 0; e < 8; e++) l2_load(bufB + 8 * t + (e ^ lo3), x[e]);
        cd16 k0r[8], k1r[8];                   // this thread's 16 key values
#if F64L2_KEYREG
        mbar_wait(kfull, (uint32_t)(i & 1));   // this step's key has landed (it was requested an inverse transform ago)
#pragma unroll
        for (int e = 0; e < 8; e++) { k0r[e] = kp[e * 64]; k1r[e] = kp[F64_CHUNK_ELEMS + e * 64]; }
        l2_fwd_pass23(x, wf3);
        bar_sync(bar_id, 64);                  // the pair has read buffer B: both buffers take the products now
#else
        l2_fwd_pass23(x, wf3);
        mbar_wait(kfull, (uint32_t)(i & 1));
        bar_sync(bar_id, 64);
#pragma unroll
        for (int e = 0; e < 8; e++) { k0r[e] = kp[e * 64]; k1r[e] = kp[F64_CHUNK_ELEMS + e * 64]; }
#endif
        {
            cd16* po = prod + (size_t)(2 * pair) * 512 + t;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const cd16 k0 = k0r[e], k1 = k1r[e];
                cd16 q0, q1;
                q0.re = F_FMA(x[e].re, k0.re, -F_MUL(x[e].im, k0.im)); q0.im = F_FMA(x[e].re, k0.im, F_MUL(x[e].im, k0.re));
                q1.re = F_FMA(x[e].re, k1.re, -F_MUL(x[e].im, k1.im)); q1.im = F_FMA(x[e].re, k1.im, F_MUL(x[e].im, k1.re));
                po[e * 64] = q0;
                po[512 + e * 64] = q1;
            }
        }
        __syncthreads();   // the twelve products are complete, the key buffer is free
        if (threadIdx.x == 0 && i + 1 < nsteps) bulk_fetch(keybuf, key + (size_t)(i + 1) * F64_STEP_ELEMS, F64L2_KEY_BYTES, kfull);
        if (pair < 2) {    // output polynomial `pair`
            cd16 wi2[4], wi3[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { wi2[k] = ti2[k * 8 + lo3]; wi3[k] = ti3[k * 64 + t]; }
            cd y[8];
            const cd16* pi = prod + (size_t)pair * 512 + t;
#pragma unroll
            for (int e = 0; e < 8; e++) l2_load(pi + e * 64, y[e]);
#pragma unroll
            for (int j = 1; j < F64L2_PAIRS; j++)
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    const cd16 v = pi[(size_t)(2 * j) * 512 + e * 64];
                    y[e].re = F_ADD(y[e].re, v.re); y[e].im = F_ADD(y[e].im, v.im);
                }
            l2_inv_pass1(y);
            bar_sync(7, 128);   // pairs 0 and 1 have both read the products: the buffers are transpose scratch again
#pragma unroll
            for (int e = 0; e < 8; e++) l2_store(bufB + 8 * t + (e ^ lo3), y[e]);
            bar_sync(bar_id, 64);
#pragma unroll
            for (int m = 0; m < 8; m++) l2_load(bufB + 64 * hi3 + 8 * m + (lo3 ^ m), y[m]);
            l2_inv_pass23(y, wi2);
#pragma unroll
            for (int m = 0; m < 8; m++) l2_store(bufA + 64 * hi3 + 8 * m + lo3, y[m]);
            bar_sync(bar_id, 64);
#pragma unroll
            for (int e = 0; e < 8; e++) l2_load(bufA + t + 64 * e, y[e]);
            l2_inv_pass23(y, wi3);
            uint32_t* ao = acc + pair * 1024;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const cd16 u = tut[e * 64 + t];
                const double zr = F_FMA(y[e].re, u.re, -F_MUL(y[e].im, u.im));
                const double zi = F_FMA(y[e].re, u.im, F_MUL(y[e].im, u.re));
                l2_red_add(ao + t + 64 * e, f64_low_word(F_ADD(zr, F64_ROUND_MAGIC)));
                l2_red_add(ao + 512 + t + 64 * e, f64_low_word(F_ADD(zi, F64_ROUND_MAGIC)));
            }
        }
        __syncthreads();   // acc is complete before the next step's rotated reads
    }

    // ---- epilogue: sample_extract_index(0) (trlwe.rs:110-121) + key-switch digits (tlwe.rs:47-64) ----
    if (a.trlwe_out) {
        uint32_t* dst = a.trlwe_out + (size_t)gate * 2048;
        for (int k = threadIdx.x; k < 2048; k += F64L2_THREADS) dst[k] = acc[k];
    }
    if (a.ksdig || a.lwe1_out) {
        for (int i = threadIdx.x; i < 1024; i += F64L2_THREADS) {
            const uint32_t ai = (i == 0) ? acc[1024] : 0u - acc[1024 + 1024 - i];
            if (a.ksdig) a.ksdig[(size_t)gate * 1024 + i] = (uint16_t)((ai + 0x8000u) >> 16);
            if (a.lwe1_out) a.lwe1_out[(size_t)gate * 1025 + 1 + i] = ai;
        }
        if (a.lwe1_out && threadIdx.x == 0) a.lwe1_out[(size_t)gate * 1025] = acc[0];
    }
    if (a.out_init) {
        uint32_t* dst = a.out_init + (size_t)(a.idxo ? (long)a.idxo[gate] : gate) * (LWE_N + 1);
        for (int c = threadIdx.x; c <= LWE_N; c += F64L2_THREADS) dst[c] = (c == 0) ? acc[0] : 0u;
    }
}

// the transformed key, polynomial by polynomial, from the one-warp layout [register 16][lane 32] (position p at (p & 15) * 32 + (p >> 4))
// to the two-warp layout [register 8][thread 64] (position p = 8 t + e at e * 64 + t)
__global__ void f64l2_key_layout_kernel(const cd16* __restrict__ src, cd16* __restrict__ dst, long npolys) {
    const long poly = blockIdx.x;
    if (poly >= npolys) return;
    for (int k = threadIdx.x; k < 512; k += blockDim.x) {
        const int e = k >> 6, t = k & 63, p = 8 * t + e;
        dst[poly * 512 + k] = src[poly * 512 + (p & 15) * 32 + (p >> 4)];
    }
}

// =====================================================================================================
// K5FL3: the same step on a CLUSTER OF TWO SMs (batches of at most #SMs/2 gates).  K5FL2 is bound by the shared-memory bandwidth
// of its SM; here CTA c owns accumulator polynomial c: its three warp pairs transform the three digits of that polynomial (half of
// the transposes, products and key traffic per SM), multiply by the key of BOTH output polynomials, and the CTA then
//   * adds its three products of output (1 - c) and sends that partial sum (8 KB) to the peer with ONE bulk DSMEM copy that counts
//     its bytes on the peer's mbarrier (warp pair 1),
//   * adds its three products of output c, waits for the peer's partial sum, adds it, runs the inverse transform of output c and
//     updates ITS accumulator polynomial (warp pair 0).
// Send and receive buffers are double buffered by step parity: the peer sends step i only after it has finished step i - 1, for
// which it needed my step i - 1, which I sent after consuming its step i - 2 -- no reverse signal is needed (the argument of the
// NTT cluster kernel, blind_rotate.cuh).  One cluster barrier at the start, one at the end.
// =====================================================================================================
constexpr int F64L3_THREADS = 3 * 64;
constexpr int F64L3_KEY_BYTES = 6 * 512 * 16;   // this CTA's three rows, both outputs
constexpr int F64L3_SMEM_BYTES = F64L2_TAB_ELEMS * 16 + 1024 * 4 /*own accumulator polynomial*/ + 3 * 2 * 512 * 16 /*transpose buffers = products*/ +
                                 F64L3_KEY_BYTES + 2 * 512 * 16 /*send*/ + 2 * 512 * 16 /*recv*/ + 640 * 2 /*abar*/ + 32 /*mbarriers*/;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(F64L3_THREADS, 1) blind_rotate_f64_latency3_kernel(const BrArgs a, const cd16* __restrict__ key) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();   // accumulator / output polynomial of this CTA
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd16* tab = reinterpret_cast<cd16*>(smem_raw);
    const cd16* tf2 = tab;
    const cd16* tf3 = tab + 32;
    const cd16* ti2 = tab + 288;
    const cd16* ti3 = tab + 320;
    const cd16* tut = tab + 576;
    uint32_t* acc = reinterpret_cast<uint32_t*>(tab + F64L2_TAB_ELEMS);   // polynomial c only
    cd16* scratch = reinterpret_cast<cd16*>(acc + 1024);
    cd16* prod = scratch;                         // [digit dw][output o][register][thread]
    cd16* keybuf = scratch + 3 * 2 * 512;         // [digit dw][output o][register 8][thread 64]
    cd16* sendb = keybuf + 6 * 512;               // [parity][512]
    cd16* recvb = sendb + 2 * 512;                // [parity][512]
    uint16_t* abar = reinterpret_cast<uint16_t*>(recvb + 2 * 512);
    uint64_t* kfull = reinterpret_cast<uint64_t*>(abar + 640);
    uint64_t* pfull = kfull + 1;                  // [parity]
    const int pair = threadIdx.x >> 6, t = threadIdx.x & 63;   // pair = digit dw of polynomial c
    cd16* bufA = scratch + (size_t)pair * 1024;
    cd16* bufB = bufA + 512;
    const int bar_id = 1 + pair;
    const long gate = blockIdx.x >> 1;
    const int nsteps = a.nsteps;
    {
        double* d = reinterpret_cast<double*>(tab);
        for (int k = threadIdx.x; k < 64; k += blockDim.x) { d[k] = g_l2_fwd2[k]; d[2 * 288 + k] = g_l2_inv2[k]; }
        for (int k = threadIdx.x; k < 512; k += blockDim.x) { d[2 * 32 + k] = g_l2_fwd3[k]; d[2 * 320 + k] = g_l2_inv3[k]; }
        for (int k = threadIdx.x; k < 1024; k += blockDim.x) d[2 * 576 + k] = g_l2_untw[k];
    }
    // ---- prologue (both CTAs): gate pre-combination (tfhe.rs:27-71), rounding of (b, a) (tfhe.rs:97,107-108), acc_0 ----
    {
        uint32_t* lin = reinterpret_cast<uint32_t*>(scratch);
        const bool second = gate >= a.split;
        const long gsrc = second ? gate - a.split : gate;
        const uint32_t* q0 = second ? a.in0b : a.in0;
        const uint32_t* q1 = second ? a.in1b : a.in1;
        uint32_t k0 = (uint32_t)(second ? a.c0b : a.c0), k1 = (uint32_t)(second ? a.c1b : a.c1), kb = second ? a.cbb : a.cb;
        if (a.ops) gate_coeffs(a.ops[gate], a.mu, k0, k1, kb);
        const uint32_t* p0 = q0 + (size_t)(a.idx0 ? (long)a.idx0[gate] : gsrc) * (LWE_N + 1);
        const uint32_t* p1 = (q1 && k1 != 0) ? q1 + (size_t)(a.idx1 ? (long)a.idx1[gate] : gsrc) * (LWE_N + 1) : nullptr;
        for (int k = threadIdx.x; k <= LWE_N; k += F64L3_THREADS) {
            uint32_t v = k0 * p0[k];
            if (p1) v += k1 * p1[k];
            if (k == 0) v += kb;
            lin[k] = v;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 640; i += F64L3_THREADS) abar[i] = i < LWE_N ? (uint16_t)((lin[1 + i] + (1u << 20)) >> 21) : (uint16_t)0;
        const uint32_t bbar = lin[0] >> 21;
        const uint32_t nrot = (2048u - bbar) & 2047u;
        __syncthreads();   // lin (the scratch) has been read
        for (int k = threadIdx.x; k < 1024; k += F64L3_THREADS) {
            const bool neg = ((uint32_t)k < (nrot & 1023u)) != (nrot >= 1024u);
            acc[k] = c == 0 ? (neg ? 0u - a.mu : a.mu) : 0u;
        }
    }
    const cd16* mykey = key + (size_t)(6 * c) * F64_CHUNK_ELEMS;   // rows 3 c .. 3 c + 2 of every step: 6 consecutive chunks
    if (threadIdx.x == 0) {
        mbar_init(kfull, 1);
        mbar_init(pfull, 1);
        mbar_init(pfull + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (nsteps > 0) bulk_fetch(keybuf, mykey, F64L3_KEY_BYTES, kfull);
    }
    __syncthreads();
    cluster.sync();   // both CTAs are set up (mbarriers) before the first remote copy

    const uint32_t remote_recv = map_to_cta(smem_u32(recvb), (uint32_t)(c ^ 1));
    const uint32_t remote_bar = map_to_cta(smem_u32(pfull), (uint32_t)(c ^ 1));
    const int hi3 = t >> 3, lo3 = t & 7;
    cd16 wf2[4], wf3[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { wf2[k] = tf2[k * 8 + hi3]; wf3[k] = tf3[k * 64 + t]; }
    const cd16* kp = keybuf + (size_t)(2 * pair) * F64_CHUNK_ELEMS + t;
    const int sh = 6 * pair;

#pragma unroll 1
    for (int i = 0; i < nsteps; i++) {
        if (threadIdx.x == 0) mbar_expect_tx(pfull + (i & 1), 512u * 16u);   // arm this step's arrival of the peer's partial sum
        cd x[8];
        {
            const uint32_t ab = abar[i];
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const uint32_t ur = add_alu(rot_diff(acc, (uint32_t)(t + 64 * e), ab), a.mask) ^ a.mask;
                const uint32_t ui = add_alu(rot_diff(acc, (uint32_t)(512 + t + 64 * e), ab), a.mask) ^ a.mask;
                x[e].re = (double)((((int32_t)(ur << sh)) >> 24) & ~3);
                x[e].im = (double)((((int32_t)(ui << sh)) >> 24) & ~3);
            }
        }
        l2_fwd_pass1(x);
#pragma unroll
        for (int e = 0; e < 8; e++) l2_store(bufA + t + 64 * e, x[e]);
        bar_sync(bar_id, 64);
#pragma unroll
        for (int m = 0; m < 8; m++) l2_load(bufA + 64 * hi3 + 8 * m + lo3, x[m]);
        l2_fwd_pass23(x, wf2);
#pragma unroll
        for (int m = 0; m < 8; m++) l2_store(bufB + 64 * hi3 + 8 * m + (lo3 ^ m), x[m]);
        bar_sync(bar_id, 64);
#pragma unroll
        for (int e = 0; e < 8; e++) l2_load(bufB + 8 * t + (e ^ lo3), x[e]);
        mbar_wait(kfull, (uint32_t)(i & 1));
        cd16 k0r[8], k1r[8];
#pragma unroll
        for (int e = 0; e < 8; e++) { k0r[e] = kp[e * 64]; k1r[e] = kp[F64_CHUNK_ELEMS + e * 64]; }
        l2_fwd_pass23(x, wf3);
        bar_sync(bar_id, 64);   // the pair has read buffer B: both buffers take the products now
        {
            cd16* po = prod + (size_t)(2 * pair) * 512 + t;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const cd16 k0 = k0r[e], k1 = k1r[e];
                cd16 q0, q1;
                q0.re = F_FMA(x[e].re, k0.re, -F_MUL(x[e].im, k0.im)); q0.im = F_FMA(x[e].re, k0.im, F_MUL(x[e].im, k0.re));
                q1.re = F_FMA(x[e].re, k1.re, -F_MUL(x[e].im, k1.im)); q1.im = F_FMA(x[e].re, k1.im, F_MUL(x[e].im, k1.re));
                po[e * 64] = q0;
                po[512 + e * 64] = q1;
            }
        }
        __syncthreads();   // the six products are complete, the key buffer is free
        if (threadIdx.x == 0 && i + 1 < nsteps) bulk_fetch(keybuf, mykey + (size_t)(i + 1) * F64_STEP_ELEMS, F64L3_KEY_BYTES, kfull);
        if (pair == 1) {   // the three products of the PEER's output polynomial: add, send
            cd16* sb = sendb + (size_t)(i & 1) * 512;
            const cd16* pi = prod + (size_t)(c ^ 1) * 512 + t;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                cd16 v = pi[e * 64];
                const cd16 v1 = pi[(size_t)2 * 512 + e * 64], v2 = pi[(size_t)4 * 512 + e * 64];
                v.re = F_ADD(F_ADD(v.re, v1.re), v2.re); v.im = F_ADD(F_ADD(v.im, v1.im), v2.im);
                sb[e * 64 + t] = v;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the sums are visible to the copy engine
            bar_sync(7, 128);   // ... of every thread of the pair; and pairs 0 and 1 have both read the products
            if (t == 0) bulk_send(remote_recv + (uint32_t)(i & 1) * 512u * 16u, sb, 512u * 16u, remote_bar + 8u * (uint32_t)(i & 1));
        } else if (pair == 0) {   // the three products of MY output polynomial, the peer's partial sum, the inverse transform
            cd16 wi2[4], wi3[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { wi2[k] = ti2[k * 8 + lo3]; wi3[k] = ti3[k * 64 + t]; }
            cd y[8];
            const cd16* pi = prod + (size_t)c * 512 + t;
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const cd16 v0 = pi[e * 64], v1 = pi[(size_t)2 * 512 + e * 64], v2 = pi[(size_t)4 * 512 + e * 64];
                y[e].re = F_ADD(F_ADD(v0.re, v1.re), v2.re); y[e].im = F_ADD(F_ADD(v0.im, v1.im), v2.im);
            }
            bar_sync(7, 128);   // pairs 0 and 1 have both read the products: the buffers are transpose scratch again
            mbar_wait(pfull + (i & 1), (uint32_t)(i >> 1) & 1u);
            {
                const cd16* rb = recvb + (size_t)(i & 1) * 512 + t;
#pragma unroll
                for (int e = 0; e < 8; e++) { const cd16 v = rb[e * 64]; y[e].re = F_ADD(y[e].re, v.re); y[e].im = F_ADD(y[e].im, v.im); }
            }
            l2_inv_pass1(y);
#pragma unroll
            for (int e = 0; e < 8; e++) l2_store(bufB + 8 * t + (e ^ lo3), y[e]);
            bar_sync(bar_id, 64);
#pragma unroll
            for (int m = 0; m < 8; m++) l2_load(bufB + 64 * hi3 + 8 * m + (lo3 ^ m), y[m]);
            l2_inv_pass23(y, wi2);
#pragma unroll
            for (int m = 0; m < 8; m++) l2_store(bufA + 64 * hi3 + 8 * m + lo3, y[m]);
            bar_sync(bar_id, 64);
#pragma unroll
            for (int e = 0; e < 8; e++) l2_load(bufA + t + 64 * e, y[e]);
            l2_inv_pass23(y, wi3);
#pragma unroll
            for (int e = 0; e < 8; e++) {
                const cd16 u = tut[e * 64 + t];
                const double zr = F_FMA(y[e].re, u.re, -F_MUL(y[e].im, u.im));
                const double zi = F_FMA(y[e].re, u.im, F_MUL(y[e].im, u.re));
                l2_red_add(acc + t + 64 * e, f64_low_word(F_ADD(zr, F64_ROUND_MAGIC)));
                l2_red_add(acc + 512 + t + 64 * e, f64_low_word(F_ADD(zi, F64_ROUND_MAGIC)));
            }
        }
        __syncthreads();   // my accumulator polynomial is complete before the next step's rotated reads
    }
    cluster.sync();   // no CTA leaves while its peer may still copy into its shared memory

    // ---- epilogue: sample_extract_index(0) (trlwe.rs:110-121) + key-switch digits (tlwe.rs:47-64): CTA 0 has b, CTA 1 has a ----
    if (a.trlwe_out) {
        uint32_t* dst = a.trlwe_out + (size_t)gate * 2048 + (size_t)c * 1024;
        for (int k = threadIdx.x; k < 1024; k += F64L3_THREADS) dst[k] = acc[k];
    }
    if (c == 1 && (a.ksdig || a.lwe1_out)) {
        for (int i = threadIdx.x; i < 1024; i += F64L3_THREADS) {
            const uint32_t ai = (i == 0) ? acc[0] : 0u - acc[1024 - i];
            if (a.ksdig) a.ksdig[(size_t)gate * 1024 + i] = (uint16_t)((ai + 0x8000u) >> 16);
            if (a.lwe1_out) a.lwe1_out[(size_t)gate * 1025 + 1 + i] = ai;
        }
    }
    if (c == 0) {
        if (a.lwe1_out && threadIdx.x == 0) a.lwe1_out[(size_t)gate * 1025] = acc[0];
        if (a.out_init) {
            uint32_t* dst = a.out_init + (size_t)(a.idxo ? (long)a.idxo[gate] : gate) * (LWE_N + 1);
            for (int k = threadIdx.x; k <= LWE_N; k += F64L3_THREADS) dst[k] = (k == 0) ? acc[0] : 0u;
        }
    }
}
