// engine.cu -- CUDA kernels (sm_100a) and C-ABI of the B200 TFHE gate-bootstrapping engine.
//
// Kernels (DESIGN.md has the roofline of each):
//   K8 bk_transform_kernel : torus-domain bootstrapping key -> NTT domain, 3 centred 11-bit slices per polynomial
//                            (replaces TRGSWRepF::from, hom_nand/src/trgsw.rs:68-76)
//   K5 blind_rotate_kernel : gate pre-combination + 635 x CMUX + sample extract (tfhe.rs:27-113, trgsw.rs:264-322,
//                            trlwe.rs:110-121), persistent per gate, state resident in shared memory
//   K6 keyswitch_kernel    : identity_key_switch as a tiled gather-accumulate (tlwe.rs:43-73)
//   polymul_kernel         : exact negacyclic product micro-entry (math.rs:337-347)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/tfhe_b200.h"
#include "cmux_steps.cuh"

using namespace tfhe;

// =====================================================================================================
// K8: key transform.  One warp per (step i, row j, poly); loops over the three slices.
// =====================================================================================================
constexpr int KT_WARPS = 4;
__global__ void __launch_bounds__(KT_WARPS * 32) bk_transform_kernel(const uint32_t* __restrict__ bk, uint32_t* __restrict__ dev,
                                                                    int npolys /* = nsteps*12 */) {
    __shared__ __align__(16) uint32_t twF[32 * TWB_STRIDE];
    __shared__ __align__(16) uint32_t scratch[KT_WARPS][1024];
    for (int t = threadIdx.x; t < 32 * TWB_STRIDE; t += blockDim.x) twF[t] = g_fwdB[t];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pid = blockIdx.x * KT_WARPS + warp;
    if (pid >= npolys) return;
    const int poly = pid & 1, j = (pid >> 1) % BK_ROWS, i = pid / (2 * BK_ROWS);
    const uint32_t* src = bk + (size_t)pid * 1024;
    uint32_t* S = scratch[warp];
    for (int part = 0; part < 3; part++) {
        key_cols(lane, src, part, S);
        __syncwarp();
        key_rows(lane, S, twF, dev + bk_off(i, poly, part, j, 0, 0));
        __syncwarp();
    }
}

// =====================================================================================================
// K5: blind rotation.  G gates per CTA, 6 warps per gate.
// =====================================================================================================
constexpr int WARPS_PER_GATE = 6;
constexpr int THREADS_PER_GATE = WARPS_PER_GATE * 32;
constexpr int GATE_SMEM_WORDS = 2 * 1024 /*acc*/ + 6 * 1024 /*dh: digit spectra / transpose scratch*/ + 320 /*abar u16[640]*/;
constexpr int TW_SMEM_WORDS = 2 * 32 * TWB_STRIDE;
constexpr size_t br_smem_bytes(int G) { return (size_t)(TW_SMEM_WORDS + G * GATE_SMEM_WORDS) * 4; }

struct BrArgs {
    const uint32_t* bkdev;   // NTT-domain key, BK_STEP_WORDS per step
    const uint32_t* in0;     // [B][n+1]
    const uint32_t* in1;     // [B][n+1] or null
    int32_t c0, c1;          // lin = c0*in0 + c1*in1 + (cb, 0, ...)
    uint32_t cb;
    uint32_t mu, mask;
    int nsteps;
    long B;
    // outputs (any may be null)
    uint32_t* out_init;      // [B][n+1]  <- (b', 0, ..., 0)  : accumulator the key-switch kernel subtracts from
    uint16_t* ksdig;         // [B][N]    <- packed key-switch digits of the extracted sample
    uint32_t* trlwe_out;     // [B][2][N]
    uint32_t* lwe1_out;      // [B][N+1]
    // external-product mode
    const uint32_t* trlwe_in;  // [B][2][N]
    long ntrgsw;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_sync(int id, int nthreads) {
#if !defined(TFHE_EXP_NOBAR)   /* timing experiment only */
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
#endif
}
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
#if defined(TFHE_EXP_NOBAR)
    return;
#endif
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

template <int G, bool EXTPROD, int MINB>
__global__ void __launch_bounds__(G* THREADS_PER_GATE, MINB) blind_rotate_kernel(const BrArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* twF = smem;
    uint32_t* twI = smem + 32 * TWB_STRIDE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gl = warp / WARPS_PER_GATE, w6 = warp % WARPS_PER_GATE;
    const int pw = w6 / 3, kw = w6 % 3;
    const int tid6 = threadIdx.x - gl * THREADS_PER_GATE;
    uint32_t* acc = smem + TW_SMEM_WORDS + gl * GATE_SMEM_WORDS;
    uint32_t* dh = acc + 2 * 1024;
    uint16_t* abar = reinterpret_cast<uint16_t*>(dh + 6 * 1024);
    uint64_t* macdone = reinterpret_cast<uint64_t*>(dh + 6 * 1024 + 318);  // abar uses 635 u16 = 317.5 words of its 320

    const long gate_raw = (long)blockIdx.x * G + gl;
    const bool active = gate_raw < a.B;
    const long gate = active ? gate_raw : a.B - 1;

    for (int t = threadIdx.x; t < 32 * TWB_STRIDE; t += blockDim.x) {
        twF[t] = g_fwdB[t];
        twI[t] = g_invB[t];
    }
    if (tid6 == 0) mbar_init(macdone, WARPS_PER_GATE);
    // ---- prologue: gate pre-combination (tfhe.rs:27-71), rounding of (b, a) (tfhe.rs:97,107-108), acc_0 ----
    int nsteps = a.nsteps;
    if (EXTPROD) {
        nsteps = 1;
        const uint32_t* src = a.trlwe_in + (size_t)gate * 2048;
        for (int k = tid6; k < 2048; k += THREADS_PER_GATE) acc[k] = src[k];
    } else {
        uint32_t* lin = dh;
        const uint32_t* p0 = a.in0 + (size_t)gate * (LWE_N + 1);
        const uint32_t* p1 = a.in1 ? a.in1 + (size_t)gate * (LWE_N + 1) : nullptr;
        for (int c = tid6; c <= LWE_N; c += THREADS_PER_GATE) {
            uint32_t v = (uint32_t)a.c0 * p0[c];
            if (p1) v += (uint32_t)a.c1 * p1[c];
            if (c == 0) v += a.cb;
            lin[c] = v;
        }
        __syncthreads();
        for (int i = tid6; i < LWE_N; i += THREADS_PER_GATE) abar[i] = (uint16_t)((lin[1 + i] + (1u << 20)) >> 21);  // round
        const uint32_t bbar = lin[0] >> 21;                                                                         // floor
        const uint32_t nrot = (2048u - bbar) & 2047u;  // acc_0 = X^{-bbar} * (mu, ..., mu ; 0)
        for (int k = tid6; k < 1024; k += THREADS_PER_GATE) {
            const bool neg = ((uint32_t)k < (nrot & 1023u)) != (nrot >= 1024u);
            acc[k] = neg ? 0u - a.mu : a.mu;
            acc[1024 + k] = 0;
        }
    }
    __syncthreads();

    // ---- 635 x CMUX ----
    // Synchronisation per step (named barriers, so gates sharing a CTA and the two polynomials of a gate decouple):
    //   B1 gate barrier (192 threads)  : the 6 digit spectra of this step are complete
    //   macdone (mbarrier, 6 arrivals) : every warp finished READING the digit spectra dh[] -> a warp may reuse its own
    //                                    plane dh[w6] as the transpose scratch of its inverse transform
    //   3 poly barriers (96 threads)   : the three key-slice warps of a polynomial add their exact slice into acc[poly]
    //                                    one after the other (no separate output planes: 24 KB less shared memory per gate)
    const int bar_gate = 1 + gl, bar_poly = 1 + G + 2 * gl + pw;
    uint32_t mac_parity = 0;
#pragma unroll 1
    for (int i = 0; i < nsteps; i++) {
        const uint32_t* step_bk = a.bkdev + (EXTPROD ? (size_t)(gate % a.ntrgsw) : (size_t)i) * BK_STEP_WORDS;
        uint32_t* S = dh + w6 * 1024;
        {   // phase 1: digit kw of poly pw -> spectrum plane dh[w6]
            p1a<!EXTPROD>(lane, acc + pw * 1024, EXTPROD ? 0u : (uint32_t)abar[i], a.mask, kw, S);
            __syncwarp();
            p1b(lane, S, twF);
        }
        bar_sync(bar_gate, THREADS_PER_GATE);
        uint32_t x[32];
        {   // phase 2: key slice kw of output poly pw
            p2a_mac(lane, step_bk + (size_t)(pw * 3 + kw) * BK_SLAB_WORDS, dh, x);
            __syncwarp();
            if (lane == 0) mbar_arrive(macdone);
            gs32(x, TwRow{twI + lane * TWB_STRIDE});
            mbar_wait(macdone, mac_parity);
            mac_parity ^= 1u;
#pragma unroll
            for (int q = 0; q < 8; q++)
                *reinterpret_cast<uint4*>(S + swz_chunk(lane, q)) = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
            __syncwarp();
            p2b(lane, S, kw, x);   // x[r] = exact slice value (already shifted) of coefficient 32 r + lane
        }
        // phase 3: acc[pw] += x, slice warps take turns (EXTPROD: the first turn overwrites)
        uint32_t* A = acc + pw * 1024 + lane;
#pragma unroll
        for (int turn = 0; turn < 3; turn++) {
            if (kw == turn) {
#pragma unroll
                for (int r = 0; r < 32; r++) A[32 * r] = (EXTPROD && turn == 0) ? x[r] : A[32 * r] + x[r];
            }
            bar_sync(bar_poly, 96);
        }
    }
    bar_sync(bar_gate, THREADS_PER_GATE);

    // ---- epilogue: sample_extract_index(0) (trlwe.rs:110-121) + key-switch digits (tlwe.rs:47-64) ----
    if (!active) return;
    if (a.trlwe_out) {
        uint32_t* dst = a.trlwe_out + (size_t)gate * 2048;
        for (int k = tid6; k < 2048; k += THREADS_PER_GATE) dst[k] = acc[k];
    }
    if (a.ksdig || a.lwe1_out) {
        for (int i = tid6; i < 1024; i += THREADS_PER_GATE) {
            const uint32_t ai = (i == 0) ? acc[1024] : 0u - acc[1024 + 1024 - i];
            if (a.ksdig) a.ksdig[(size_t)gate * 1024 + i] = (uint16_t)((ai + 0x8000u) >> 16);
            if (a.lwe1_out) a.lwe1_out[(size_t)gate * 1025 + 1 + i] = ai;
        }
        if (a.lwe1_out && tid6 == 0) a.lwe1_out[(size_t)gate * 1025] = acc[0];
    }
    if (a.out_init) {
        uint32_t* dst = a.out_init + (size_t)gate * (LWE_N + 1);
        for (int c = tid6; c <= LWE_N; c += THREADS_PER_GATE) dst[c] = (c == 0) ? acc[0] : 0u;
    }
}

// =====================================================================================================
// K6: key switch.  out[g] -= sum_{i,l : d != 0} KSK[i][l][d-1]  with d = 2-bit digit (i,l) of gate g.
// CTA = (tile of KS_GT gates) x (slice of 1024/KS_ISPLIT key indices); thread = one 16-byte column chunk of the
// 636-word rows (159 chunks).  Every KSK row is read once per CTA and applied to all gates of the tile; the digit
// is CTA-uniform so the select is a uniform branch.  Partial sums are merged with red.global.add.u32.
// =====================================================================================================
constexpr int KS_GT = 16;
constexpr int KS_ISPLIT = 8;
constexpr int KS_ICHUNK = 1024 / KS_ISPLIT;
constexpr int KS_CHUNKS = (LWE_N + 1) / 4;  // 159
constexpr int KS_THREADS = 160;
__global__ void __launch_bounds__(KS_THREADS) keyswitch_kernel(const uint4* __restrict__ ksk, const uint16_t* __restrict__ dig,
                                                              uint32_t* __restrict__ out, long B) {
    __shared__ __align__(16) uint16_t dg[KS_ICHUNK][KS_GT];
    const long g0 = (long)blockIdx.x * KS_GT;
    const int i0 = blockIdx.y * KS_ICHUNK;
    for (int t = threadIdx.x; t < KS_ICHUNK * KS_GT; t += blockDim.x) {
        const int g = t / KS_ICHUNK, ii = t % KS_ICHUNK;
        dg[ii][g] = (g0 + g < B) ? dig[(size_t)(g0 + g) * 1024 + i0 + ii] : (uint16_t)0;
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t >= KS_CHUNKS) return;
    uint4 acc[KS_GT];
#pragma unroll
    for (int g = 0; g < KS_GT; g++) acc[g] = make_uint4(0, 0, 0, 0);
    const uint4* base = ksk + (size_t)i0 * 8 * 3 * KS_CHUNKS + t;
#pragma unroll 1
    for (int ii = 0; ii < KS_ICHUNK; ii++) {
        uint32_t dw[KS_GT / 2];
#pragma unroll
        for (int g = 0; g < KS_GT / 2; g++) dw[g] = reinterpret_cast<const uint32_t*>(dg[ii])[g];  // two gates per word
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
        uint32_t* o = out + (size_t)(g0 + g) * (LWE_N + 1) + 4 * t;
        atomicAdd(o + 0, 0u - acc[g].x);
        atomicAdd(o + 1, 0u - acc[g].y);
        atomicAdd(o + 2, 0u - acc[g].z);
        atomicAdd(o + 3, 0u - acc[g].w);
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

// =====================================================================================================
// exact negacyclic product a (torus) * d (small ints): one warp per product, 7 transforms
// =====================================================================================================
constexpr int PM_WARPS = 2;
__global__ void __launch_bounds__(PM_WARPS * 32) polymul_kernel(const uint32_t* __restrict__ A, const int32_t* __restrict__ D,
                                                               uint32_t* __restrict__ out, long B) {
    __shared__ __align__(16) uint32_t twF[32 * TWB_STRIDE];
    __shared__ __align__(16) uint32_t twI[32 * TWB_STRIDE];
    __shared__ __align__(16) uint32_t scratch[PM_WARPS][2][1024];
    for (int t = threadIdx.x; t < 32 * TWB_STRIDE; t += blockDim.x) { twF[t] = g_fwdB[t]; twI[t] = g_invB[t]; }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long g = (long)blockIdx.x * PM_WARPS + warp;
    if (g >= B) return;
    const uint32_t* a = A + (size_t)g * 1024;
    const int32_t* d = D + (size_t)g * 1024;
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
#pragma unroll
    for (int r = 0; r < 32; r++) out[(size_t)g * 1024 + 32 * r + lane] = res[r];
}

// =====================================================================================================
// host side: context + C ABI
// =====================================================================================================
struct tfhe_b200_ctx {
    tfhe_b200_params prm;
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    uint32_t* bkdev = nullptr;   // n * BK_STEP_WORDS
    uint32_t* kskdev = nullptr;  // [N][t][3][n+1]
    bool have_bk = false, have_ksk = false;
    // workspaces (grown on demand)
    uint16_t* ksdig = nullptr; size_t ksdig_cap = 0;
    uint32_t* tmp[4] = {nullptr, nullptr, nullptr, nullptr}; size_t tmp_cap[4] = {0, 0, 0, 0};
    uint32_t* xbk = nullptr; size_t xbk_cap = 0;  // external-product scratch keys
    static constexpr int RING = 64;          // event ring: per-launch device times of the last RING timed gate batches
    cudaEvent_t ev[RING][4] = {};
    uint64_t timed = 0;
    uint64_t launches = 0;
    uint64_t last_batch = 0;
    int gates_per_cta = 2;
    int variant = 3;  // blind-rotate launch shape: 0 = 2 gates/CTA x 1 CTA/SM, 2 = 1 gate/CTA x 2 CTA/SM, 3 = 1 gate/CTA x 3 CTA/SM
    std::string err;
};
static thread_local std::string g_create_err;

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                \
            return TFHE_B200_ERR_CUDA;                                                                    \
        }                                                                                                 \
    } while (0)

static int fail(tfhe_b200_ctx* ctx, int code, const char* msg) {
    if (ctx) ctx->err = msg; else g_create_err = msg;
    return code;
}
static int grow(tfhe_b200_ctx* ctx, void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return TFHE_B200_OK;
    if (*p) CK(cudaFree(*p));
    *p = nullptr; *cap = 0;
    CK(cudaMalloc(p, bytes));
    *cap = bytes;
    return TFHE_B200_OK;
}

extern "C" {

const char* tfhe_b200_version(void) { return "rustfhe_b200 0.1 (sm_100a, p=536856577, 3x11-bit key slices)"; }

int tfhe_b200_default_params(tfhe_b200_params* p) {
    if (!p) return TFHE_B200_ERR_PARAM;
    p->n = 635; p->N = 1024; p->l = 3; p->bgbit = 6; p->ks_t = 8; p->ks_basebit = 2;
    p->mu = 0x20000000u; p->decomp_mask = TFHE_B200_MASK_FAITHFUL;
    return TFHE_B200_OK;
}

int tfhe_b200_ctx_create(const tfhe_b200_params* p, int device, tfhe_b200_ctx** out) {
    if (!out) return fail(nullptr, TFHE_B200_ERR_PARAM, "ctx_create: out is NULL");
    *out = nullptr;
    tfhe_b200_params prm;
    tfhe_b200_default_params(&prm);
    if (p) prm = *p;
    if (prm.n != 635 || prm.N != 1024 || prm.l != 3 || prm.bgbit != 6 || prm.ks_t != 8 || prm.ks_basebit != 2)
        return fail(nullptr, TFHE_B200_ERR_PARAM, "ctx_create: this build supports n=635 N=1024 l=3 Bgbit=6 t=8 basebit=2 only");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("ctx_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
        return TFHE_B200_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) return fail(nullptr, TFHE_B200_ERR_PARAM, "ctx_create: bad device index");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return TFHE_B200_ERR_CUDA; }
    if (prop.major != 10) {
        g_create_err = "ctx_create: device is not sm_100 (kernels are built for sm_100a only; no fallback)";
        return TFHE_B200_ERR_CUDA;
    }
    tfhe_b200_ctx* ctx = new tfhe_b200_ctx();
    ctx->prm = prm; ctx->device = device; ctx->sm_count = prop.multiProcessorCount;
    auto bail = [&](const char* what, cudaError_t ee) { g_create_err = std::string(what) + ": " + cudaGetErrorString(ee); delete ctx; return TFHE_B200_ERR_CUDA; };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    for (auto& slot : ctx->ev) for (auto& ev : slot) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc(&ctx->bkdev, (size_t)LWE_N * BK_STEP_WORDS * 4)) != cudaSuccess) return bail("cudaMalloc(bk)", e);
    if ((e = cudaMalloc(&ctx->kskdev, (size_t)1024 * 8 * 3 * (LWE_N + 1) * 4)) != cudaSuccess) return bail("cudaMalloc(ksk)", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_kernel<2, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)br_smem_bytes(2))) != cudaSuccess) return bail("smem attr", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_kernel<1, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)br_smem_bytes(1))) != cudaSuccess) return bail("smem attr", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_kernel<1, false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)br_smem_bytes(1))) != cudaSuccess) return bail("smem attr", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_kernel<1, false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)br_smem_bytes(1))) != cudaSuccess) return bail("smem attr", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_kernel<1, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)br_smem_bytes(1))) != cudaSuccess) return bail("smem attr", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_kernel<3, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)br_smem_bytes(3))) != cudaSuccess) return bail("smem attr", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_kernel<2, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)br_smem_bytes(2))) != cudaSuccess) return bail("smem attr", e);
    if (const char* v = getenv("TFHE_B200_BR_VARIANT")) ctx->variant = atoi(v);
    *out = ctx;
    return TFHE_B200_OK;
}

int tfhe_b200_ctx_destroy(tfhe_b200_ctx* ctx) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->bkdev); cudaFree(ctx->kskdev); cudaFree(ctx->ksdig); cudaFree(ctx->xbk);
    for (auto p : ctx->tmp) cudaFree(p);
    for (auto& slot : ctx->ev) for (auto ev : slot) if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return TFHE_B200_OK;
}

const char* tfhe_b200_last_error(const tfhe_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int tfhe_b200_set_decomp_mask(tfhe_b200_ctx* ctx, uint32_t mask) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    ctx->prm.decomp_mask = mask;
    return TFHE_B200_OK;
}

int tfhe_b200_reset_stats(tfhe_b200_ctx* ctx) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    ctx->timed = 0;
    return TFHE_B200_OK;
}
int tfhe_b200_get_stats(tfhe_b200_ctx* ctx, tfhe_b200_stats* out) {
    if (!ctx || !out) return TFHE_B200_ERR_PARAM;
    memset(out, 0, sizeof *out);
    out->kernel_launches = ctx->launches;
    out->last_batch = ctx->last_batch;
    out->gates_per_cta = ctx->gates_per_cta;
    out->sm_count = ctx->sm_count;
    out->device_key_bytes = (uint64_t)LWE_N * BK_STEP_WORDS * 4 + (uint64_t)1024 * 8 * 3 * (LWE_N + 1) * 4;
    const uint64_t cnt = ctx->timed < (uint64_t)tfhe_b200_ctx::RING ? ctx->timed : (uint64_t)tfhe_b200_ctx::RING;
    double sb = 0, sk = 0;
    for (uint64_t k = 0; k < cnt; k++) {
        const int slot = (int)((ctx->timed - 1 - k) % tfhe_b200_ctx::RING);
        float b = 0, s = 0;
        CK(cudaEventSynchronize(ctx->ev[slot][3]));
        CK(cudaEventElapsedTime(&b, ctx->ev[slot][0], ctx->ev[slot][1]));
        CK(cudaEventElapsedTime(&s, ctx->ev[slot][2], ctx->ev[slot][3]));
        if (k == 0) { out->last_blind_rotate_ms = b; out->last_keyswitch_ms = s; }
        sb += b; sk += s;
    }
    out->timed_launches = cnt;
    if (cnt) { out->avg_blind_rotate_ms = (float)(sb / cnt); out->avg_keyswitch_ms = (float)(sk / cnt); }
    return TFHE_B200_OK;
}

// ---- keys ----
static int transform_keys(tfhe_b200_ctx* ctx, const uint32_t* src_dev, uint32_t* dst_dev, int nsteps, cudaStream_t st) {
    const int npolys = nsteps * 12;
    bk_transform_kernel<<<(npolys + KT_WARPS - 1) / KT_WARPS, KT_WARPS * 32, 0, st>>>(src_dev, dst_dev, npolys);
    ctx->launches++;
    CK(cudaGetLastError());
    return TFHE_B200_OK;
}
int tfhe_b200_load_bk_device(tfhe_b200_ctx* ctx, const uint32_t* bk_dev, void* stream) {
    if (!ctx || !bk_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "load_bk_device: null argument");
    CK(cudaSetDevice(ctx->device));
    int rc = transform_keys(ctx, bk_dev, ctx->bkdev, LWE_N, (cudaStream_t)stream);
    if (rc) return rc;
    ctx->have_bk = true;
    return TFHE_B200_OK;
}
int tfhe_b200_load_bk(tfhe_b200_ctx* ctx, const uint32_t* bk_host) {
    if (!ctx || !bk_host) return fail(ctx, TFHE_B200_ERR_PARAM, "load_bk: null argument");
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)LWE_N * 12 * 1024 * 4;
    uint32_t* staging = nullptr;
    CK(cudaMalloc(&staging, bytes));
    cudaError_t e = cudaMemcpyAsync(staging, bk_host, bytes, cudaMemcpyHostToDevice, ctx->stream);
    int rc = TFHE_B200_OK;
    if (e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = TFHE_B200_ERR_CUDA; }
    if (!rc) rc = tfhe_b200_load_bk_device(ctx, staging, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    if (!rc && e != cudaSuccess) { ctx->err = cudaGetErrorString(e); rc = TFHE_B200_ERR_CUDA; }
    cudaFree(staging);
    return rc;
}
int tfhe_b200_load_ksk_device(tfhe_b200_ctx* ctx, const uint32_t* ksk_dev, void* stream) {
    if (!ctx || !ksk_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "load_ksk_device: null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->kskdev, ksk_dev, (size_t)1024 * 8 * 3 * (LWE_N + 1) * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    ctx->have_ksk = true;
    return TFHE_B200_OK;
}
int tfhe_b200_load_ksk(tfhe_b200_ctx* ctx, const uint32_t* ksk_host) {
    if (!ctx || !ksk_host) return fail(ctx, TFHE_B200_ERR_PARAM, "load_ksk: null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->kskdev, ksk_host, (size_t)1024 * 8 * 3 * (LWE_N + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_ksk = true;
    return TFHE_B200_OK;
}

// ---- launches ----
static void op_coeffs(int op, uint32_t mu, int32_t* c0, int32_t* c1, uint32_t* cb, bool* two) {
    *two = true;
    switch (op) {
    case TFHE_B200_NAND: *c0 = -1; *c1 = -1; *cb = mu; break;
    case TFHE_B200_AND: *c0 = 1; *c1 = 1; *cb = 0u - mu; break;
    case TFHE_B200_OR: *c0 = 1; *c1 = 1; *cb = mu; break;
    case TFHE_B200_XOR: *c0 = 2; *c1 = 2; *cb = 2u * mu; break;
    case TFHE_B200_NOT: *c0 = -1; *c1 = 0; *cb = 0; *two = false; break;
    case TFHE_B200_ANDNY: *c0 = -1; *c1 = 1; *cb = 0u - mu; break;
    default: *c0 = 1; *c1 = 0; *cb = 0; *two = false; break;
    }
}
static int launch_blind_rotate(tfhe_b200_ctx* ctx, BrArgs& a, cudaStream_t st, bool timed) {
    a.bkdev = ctx->bkdev; a.mask = ctx->prm.decomp_mask; a.mu = ctx->prm.mu;
    const int slot = (int)(ctx->timed % tfhe_b200_ctx::RING);
    if (timed) CK(cudaEventRecord(ctx->ev[slot][0], st));
    // one gate per CTA when the batch cannot fill the machine with pairs (latency case), else two
    const bool pair = a.B > (long)ctx->sm_count;
    if (pair && ctx->variant == 0) {
        const unsigned grid = (unsigned)((a.B + 1) / 2);
        blind_rotate_kernel<2, false, 1><<<grid, 2 * THREADS_PER_GATE, br_smem_bytes(2), st>>>(a);
        ctx->gates_per_cta = 2;
    } else if (pair && ctx->variant == 2) {
        blind_rotate_kernel<1, false, 2><<<(unsigned)a.B, THREADS_PER_GATE, br_smem_bytes(1), st>>>(a);
        ctx->gates_per_cta = 1;
    } else if (pair && ctx->variant == 4) {
        const unsigned grid = (unsigned)((a.B + 2) / 3);
        blind_rotate_kernel<3, false, 1><<<grid, 3 * THREADS_PER_GATE, br_smem_bytes(3), st>>>(a);
        ctx->gates_per_cta = 3;
    } else if (pair && ctx->variant == 5) {
        blind_rotate_kernel<1, false, 4><<<(unsigned)a.B, THREADS_PER_GATE, br_smem_bytes(1), st>>>(a);
        ctx->gates_per_cta = 1;
    } else if (pair && ctx->variant == 3) {
        blind_rotate_kernel<1, false, 3><<<(unsigned)a.B, THREADS_PER_GATE, br_smem_bytes(1), st>>>(a);
        ctx->gates_per_cta = 1;
    } else {
        blind_rotate_kernel<1, false, 1><<<(unsigned)a.B, THREADS_PER_GATE, br_smem_bytes(1), st>>>(a);
        ctx->gates_per_cta = 1;
    }
    ctx->launches++;
    CK(cudaGetLastError());
    if (timed) CK(cudaEventRecord(ctx->ev[slot][1], st));
    return TFHE_B200_OK;
}
static int launch_keyswitch(tfhe_b200_ctx* ctx, const uint16_t* dig, uint32_t* out, long B, cudaStream_t st, bool timed) {
    const int slot = (int)(ctx->timed % tfhe_b200_ctx::RING);
    if (timed) CK(cudaEventRecord(ctx->ev[slot][2], st));
    dim3 grid((unsigned)((B + KS_GT - 1) / KS_GT), KS_ISPLIT);
    keyswitch_kernel<<<grid, KS_THREADS, 0, st>>>(reinterpret_cast<const uint4*>(ctx->kskdev), dig, out, B);
    ctx->launches++;
    CK(cudaGetLastError());
    if (timed) { CK(cudaEventRecord(ctx->ev[slot][3], st)); ctx->timed++; }
    return TFHE_B200_OK;
}
static const size_t CT_BYTES = (size_t)(LWE_N + 1) * 4;

int tfhe_b200_gate_batch_device(tfhe_b200_ctx* ctx, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B,
                                void* stream) {
    if (!ctx || !in0 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch: null argument");
    if (op < 0 || op > TFHE_B200_ANDNY) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch: bad opcode");
    if (!ctx->have_bk || !ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "gate_batch: keys not loaded");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = grow(ctx, (void**)&ctx->ksdig, &ctx->ksdig_cap, B * 1024 * sizeof(uint16_t));
    if (rc) return rc;
    BrArgs a{};
    bool two;
    op_coeffs(op, ctx->prm.mu, &a.c0, &a.c1, &a.cb, &two);
    if (two && !in1) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch: in1 required for this opcode");
    a.in0 = in0; a.in1 = two ? in1 : nullptr; a.nsteps = LWE_N; a.B = (long)B;
    a.out_init = out; a.ksdig = ctx->ksdig;
    if ((rc = launch_blind_rotate(ctx, a, st, true))) return rc;
    if ((rc = launch_keyswitch(ctx, ctx->ksdig, out, (long)B, st, true))) return rc;
    ctx->last_batch = B;
    return TFHE_B200_OK;
}

// host-pointer wrapper: H2D, run, D2H on the ctx stream
static int with_host_io(tfhe_b200_ctx* ctx, const void* const* ins, const size_t* in_bytes, int nin, void* outp, size_t out_bytes,
                        int (*fn)(tfhe_b200_ctx*, uint32_t* const* dev_in, uint32_t* dev_out, void* user), void* user) {
    CK(cudaSetDevice(ctx->device));
    uint32_t* dev_in[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < nin; k++) {
        if (!ins[k]) continue;
        int rc = grow(ctx, (void**)&ctx->tmp[k], &ctx->tmp_cap[k], in_bytes[k]);
        if (rc) return rc;
        CK(cudaMemcpyAsync(ctx->tmp[k], ins[k], in_bytes[k], cudaMemcpyHostToDevice, ctx->stream));
        dev_in[k] = ctx->tmp[k];
    }
    int rc = grow(ctx, (void**)&ctx->tmp[3], &ctx->tmp_cap[3], out_bytes);
    if (rc) return rc;
    if ((rc = fn(ctx, dev_in, ctx->tmp[3], user))) return rc;
    CK(cudaMemcpyAsync(outp, ctx->tmp[3], out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return TFHE_B200_OK;
}

struct GateCall { int op; size_t B; };
int tfhe_b200_gate_batch(tfhe_b200_ctx* ctx, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B) {
    if (!ctx || !in0 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch: null argument");
    if (B == 0) return TFHE_B200_OK;
    const void* ins[2] = {in0, in1};
    const size_t nb[2] = {B * CT_BYTES, B * CT_BYTES};
    GateCall gc{op, B};
    return with_host_io(ctx, ins, nb, 2, out, B * CT_BYTES,
                        [](tfhe_b200_ctx* c, uint32_t* const* di, uint32_t* dout, void* u) {
                            GateCall* g = (GateCall*)u;
                            return tfhe_b200_gate_batch_device(c, g->op, di[0], di[1], dout, g->B, c->stream);
                        }, &gc);
}
int tfhe_b200_bootstrap_batch(tfhe_b200_ctx* ctx, const uint32_t* in, uint32_t* out, size_t B) {
    return tfhe_b200_gate_batch(ctx, TFHE_B200_COPY, in, nullptr, out, B);
}

int tfhe_b200_mux_batch_device(tfhe_b200_ctx* ctx, const uint32_t* control, const uint32_t* in0, const uint32_t* in1, uint32_t* out,
                               size_t B, void* stream) {
    if (!ctx || !control || !in0 || !in1 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "mux_batch: null argument");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    // i_1 = hom_and(control, input_1); i_0 = hom_and(-control, input_0); out = bootstrap(i_1 + i_0 + 1/8)  (tfhe.rs:27-40)
    uint32_t *t1 = nullptr, *t0 = nullptr;
    static_assert(sizeof(void*) == 8, "64-bit only");
    int rc;
    // scratch for the two intermediate batches lives in xbk-independent buffers
    if ((rc = grow(ctx, (void**)&ctx->xbk, &ctx->xbk_cap, 2 * B * CT_BYTES))) return rc;
    t1 = ctx->xbk; t0 = ctx->xbk + B * (LWE_N + 1);
    if ((rc = tfhe_b200_gate_batch_device(ctx, TFHE_B200_AND, control, in1, t1, B, stream))) return rc;
    if ((rc = tfhe_b200_gate_batch_device(ctx, TFHE_B200_ANDNY, control, in0, t0, B, stream))) return rc;
    return tfhe_b200_gate_batch_device(ctx, TFHE_B200_OR, t1, t0, out, B, stream);
}
int tfhe_b200_mux_batch(tfhe_b200_ctx* ctx, const uint32_t* control, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B) {
    if (!ctx || !control || !in0 || !in1 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "mux_batch: null argument");
    if (B == 0) return TFHE_B200_OK;
    const void* ins[3] = {control, in0, in1};
    const size_t nb[3] = {B * CT_BYTES, B * CT_BYTES, B * CT_BYTES};
    size_t Bc = B;
    return with_host_io(ctx, ins, nb, 3, out, B * CT_BYTES,
                        [](tfhe_b200_ctx* c, uint32_t* const* di, uint32_t* dout, void* u) {
                            return tfhe_b200_mux_batch_device(c, di[0], di[1], di[2], dout, *(size_t*)u, c->stream);
                        }, &Bc);
}

// ---- step-level entries ----
struct BrCall { int nsteps; size_t B; int what; };  // what: 0 = trlwe, 1 = lwe1
int tfhe_b200_blind_rotate_batch(tfhe_b200_ctx* ctx, const uint32_t* in, int nsteps, uint32_t* out_trlwe, size_t B) {
    if (!ctx || !in || !out_trlwe) return fail(ctx, TFHE_B200_ERR_PARAM, "blind_rotate_batch: null argument");
    if (nsteps < 0 || nsteps > LWE_N) return fail(ctx, TFHE_B200_ERR_PARAM, "blind_rotate_batch: nsteps out of range");
    if (!ctx->have_bk) return fail(ctx, TFHE_B200_ERR_STATE, "blind_rotate_batch: bootstrapping key not loaded");
    if (B == 0) return TFHE_B200_OK;
    const void* ins[1] = {in};
    const size_t nb[1] = {B * CT_BYTES};
    BrCall bc{nsteps, B, 0};
    return with_host_io(ctx, ins, nb, 1, out_trlwe, B * 2048 * 4,
                        [](tfhe_b200_ctx* c, uint32_t* const* di, uint32_t* dout, void* u) {
                            BrCall* b = (BrCall*)u;
                            BrArgs a{};
                            a.c0 = 1; a.in0 = di[0]; a.nsteps = b->nsteps; a.B = (long)b->B; a.trlwe_out = dout;
                            return launch_blind_rotate(c, a, c->stream, false);
                        }, &bc);
}
int tfhe_b200_bootstrap_lv1_batch(tfhe_b200_ctx* ctx, const uint32_t* in, uint32_t* out_lwe1, size_t B) {
    if (!ctx || !in || !out_lwe1) return fail(ctx, TFHE_B200_ERR_PARAM, "bootstrap_lv1_batch: null argument");
    if (!ctx->have_bk) return fail(ctx, TFHE_B200_ERR_STATE, "bootstrap_lv1_batch: bootstrapping key not loaded");
    if (B == 0) return TFHE_B200_OK;
    const void* ins[1] = {in};
    const size_t nb[1] = {B * CT_BYTES};
    BrCall bc{LWE_N, B, 1};
    return with_host_io(ctx, ins, nb, 1, out_lwe1, B * 1025 * 4,
                        [](tfhe_b200_ctx* c, uint32_t* const* di, uint32_t* dout, void* u) {
                            BrCall* b = (BrCall*)u;
                            BrArgs a{};
                            a.c0 = 1; a.in0 = di[0]; a.nsteps = b->nsteps; a.B = (long)b->B; a.lwe1_out = dout;
                            return launch_blind_rotate(c, a, c->stream, false);
                        }, &bc);
}
int tfhe_b200_keyswitch_batch(tfhe_b200_ctx* ctx, const uint32_t* lwe1, uint32_t* out, size_t B) {
    if (!ctx || !lwe1 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "keyswitch_batch: null argument");
    if (!ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "keyswitch_batch: key-switching key not loaded");
    if (B == 0) return TFHE_B200_OK;
    const void* ins[1] = {lwe1};
    const size_t nb[1] = {B * 1025 * 4};
    size_t Bc = B;
    return with_host_io(ctx, ins, nb, 1, out, B * CT_BYTES,
                        [](tfhe_b200_ctx* c, uint32_t* const* di, uint32_t* dout, void* u) {
                            const size_t Bn = *(size_t*)u;
                            int rc = grow(c, (void**)&c->ksdig, &c->ksdig_cap, Bn * 1024 * sizeof(uint16_t));
                            if (rc) return rc;
                            lwe1_prepare_kernel<<<(unsigned)Bn, 256, 0, c->stream>>>(di[0], c->ksdig, dout, (long)Bn);
                            c->launches++;
                            return launch_keyswitch(c, c->ksdig, dout, (long)Bn, c->stream, false);
                        }, &Bc);
}
struct XpCall { const uint32_t* trgsw; size_t ntrgsw; size_t B; };
int tfhe_b200_external_product_batch(tfhe_b200_ctx* ctx, const uint32_t* trgsw, size_t ntrgsw, const uint32_t* trlwe, uint32_t* out,
                                     size_t B) {
    if (!ctx || !trgsw || !trlwe || !out || ntrgsw == 0) return fail(ctx, TFHE_B200_ERR_PARAM, "external_product_batch: bad argument");
    if (B == 0) return TFHE_B200_OK;
    const void* ins[2] = {trlwe, trgsw};
    const size_t nb[2] = {B * 2048 * 4, ntrgsw * 12 * 1024 * 4};
    XpCall xc{trgsw, ntrgsw, B};
    return with_host_io(ctx, ins, nb, 2, out, B * 2048 * 4,
                        [](tfhe_b200_ctx* c, uint32_t* const* di, uint32_t* dout, void* u) {
                            XpCall* x = (XpCall*)u;
                            int rc = grow(c, (void**)&c->xbk, &c->xbk_cap, x->ntrgsw * BK_STEP_WORDS * 4);
                            if (rc) return rc;
                            if ((rc = transform_keys(c, di[1], c->xbk, (int)x->ntrgsw, c->stream))) return rc;
                            BrArgs a{};
                            a.bkdev = c->xbk; a.mask = c->prm.decomp_mask; a.mu = c->prm.mu; a.B = (long)x->B; a.nsteps = 1;
                            a.trlwe_in = di[0]; a.trlwe_out = dout; a.ntrgsw = (long)x->ntrgsw;
                            const unsigned grid = (unsigned)((x->B + 1) / 2);
                            blind_rotate_kernel<2, true, 1><<<grid, 2 * THREADS_PER_GATE, br_smem_bytes(2), c->stream>>>(a);
                            c->launches++;
                            cudaError_t e = cudaGetLastError();
                            if (e != cudaSuccess) { c->err = cudaGetErrorString(e); return TFHE_B200_ERR_CUDA; }
                            return TFHE_B200_OK;
                        }, &xc);
}
int tfhe_b200_negacyclic_mul_batch(tfhe_b200_ctx* ctx, const uint32_t* a, const int32_t* d, uint32_t* out, size_t B) {
    if (!ctx || !a || !d || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "negacyclic_mul_batch: null argument");
    if (B == 0) return TFHE_B200_OK;
    const void* ins[2] = {a, d};
    const size_t nb[2] = {B * 4096, B * 4096};
    size_t Bc = B;
    return with_host_io(ctx, ins, nb, 2, out, B * 4096,
                        [](tfhe_b200_ctx* c, uint32_t* const* di, uint32_t* dout, void* u) {
                            const size_t Bn = *(size_t*)u;
                            polymul_kernel<<<(unsigned)((Bn + PM_WARPS - 1) / PM_WARPS), PM_WARPS * 32, 0, c->stream>>>(
                                di[0], (const int32_t*)di[1], dout, (long)Bn);
                            c->launches++;
                            cudaError_t e = cudaGetLastError();
                            if (e != cudaSuccess) { c->err = cudaGetErrorString(e); return TFHE_B200_ERR_CUDA; }
                            return TFHE_B200_OK;
                        }, &Bc);
}

}  // extern "C"
