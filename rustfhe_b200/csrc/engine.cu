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
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/tfhe_b200.h"
#include "cmux_steps.cuh"
#include "tfhe_rng.cuh"

using namespace tfhe;

// =====================================================================================================
// K8: key transform.  One warp per (step i, row j, poly); loops over the three slices.
// =====================================================================================================
constexpr int KT_WARPS = 4;
__global__ void __launch_bounds__(KT_WARPS * 32) bk_transform_kernel(const uint32_t* __restrict__ bk, uint32_t* __restrict__ dev,
                                                                    int npolys /* = nsteps*12 */, int ns /* key slices: 3 or 2 */) {
    __shared__ __align__(16) uint32_t twF[32 * TWB_STRIDE];
    __shared__ __align__(16) uint32_t scratch[KT_WARPS][TILE_WORDS];
    for (int t = threadIdx.x; t < 32 * TWB_STRIDE; t += blockDim.x) twF[t] = g_fwdB[t];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pid = blockIdx.x * KT_WARPS + warp;
    if (pid >= npolys) return;
    const int poly = pid & 1, j = (pid >> 1) % BK_ROWS, i = pid / (2 * BK_ROWS);
    const uint32_t* src = bk + (size_t)pid * 1024;
    uint32_t* S = scratch[warp];
    for (int part = 0; part < ns; part++) {
        key_cols(lane, src, part, S, ns);
        __syncwarp();
        key_rows(lane, S, twF, dev + bk_off(i, poly, part, j, 0, 0, ns));
        __syncwarp();
    }
}

// =====================================================================================================
// K5: blind rotation.  G gates per CTA, 6 warps per gate.
// =====================================================================================================
constexpr int WARPS_PER_GATE = 6;
constexpr int THREADS_PER_GATE = WARPS_PER_GATE * 32;
constexpr int GATE_SMEM_WORDS = 2 * 1024 /*acc*/ + 2 * 1024 /*U: masked source polynomials*/ + 6 * TILE_WORDS /*dh: digit spectra / transpose scratch*/ +
                                320 /*abar u16[640]*/;
constexpr int TW_SMEM_WORDS = 2 * 32 * TWB_STRIDE + DIGIT_TAB_WORDS;   // forward + inverse twiddle rows, digit table
constexpr size_t br_smem_bytes(int G) { return (size_t)(TW_SMEM_WORDS + G * GATE_SMEM_WORDS) * 4; }

struct BrArgs {
    const uint32_t* bkdev;   // NTT-domain key, BK_STEP_WORDS per step
    const uint32_t* in0;     // [B][n+1]
    const uint32_t* in1;     // [B][n+1] or null
    int32_t c0, c1;          // lin = c0*in0 + c1*in1 + (cb, 0, ...)
    uint32_t cb;
    // second operand set for gates >= split (fused hom_mux first stage: two different gates in one launch); split = B when unused
    long split;
    const uint32_t* in0b;
    const uint32_t* in1b;
    int32_t c0b, c1b;
    uint32_t cbb;
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
    const uint32_t* trlwe_in0; // [B][2][N] or null: cmux, the product is taken of (trlwe_in - trlwe_in0) and trlwe_in0 is added back
    long ntrgsw;
    // gate -> CTA distribution (see the kernel prologue)
    int cta_base, cta_rem;
    int ns;                  // key slices per polynomial: 3 (exact in the worst case) or 2 (opt-in fast mode)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

template <int G, bool EXTPROD, int MINB, int NS = 3>
__global__ void __launch_bounds__(G* THREADS_PER_GATE, MINB) blind_rotate_kernel(const BrArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* twF = smem;
    uint32_t* twI = smem + 32 * TWB_STRIDE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gl = warp / WARPS_PER_GATE, w6 = warp % WARPS_PER_GATE;
    const int pw = w6 / 3, kw = w6 % 3;
    const int tid6 = threadIdx.x - gl * THREADS_PER_GATE;
    uint32_t* acc = smem + TW_SMEM_WORDS + gl * GATE_SMEM_WORDS;
    uint32_t* U = acc + 2 * 1024;
    uint32_t* dh = U + 2 * 1024;
    uint16_t* abar = reinterpret_cast<uint16_t*>(dh + 6 * TILE_WORDS);
    uint64_t* macdone = reinterpret_cast<uint64_t*>(dh + 6 * TILE_WORDS + 318);  // abar uses 635 u16 = 317.5 words of its 320

    // gates are dealt out evenly: the first cta_rem CTAs own cta_base+1 consecutive gates, the others cta_base (<= G)
    const long cta = blockIdx.x;
    const long first = cta * a.cta_base + (cta < a.cta_rem ? cta : a.cta_rem);
    const int cnt = a.cta_base + (cta < a.cta_rem ? 1 : 0);
    const bool active = gl < cnt;
    const long gate = active ? first + gl : a.B - 1;

    uint32_t* dtab = smem + 2 * 32 * TWB_STRIDE;
    for (int t = threadIdx.x; t < 32 * TWB_STRIDE; t += blockDim.x) {
        twF[t] = g_fwdB[t];
        twI[t] = g_invB[t];
    }
    for (int t = threadIdx.x; t < DIGIT_TAB_WORDS; t += blockDim.x) dtab[t] = g_digit_tab.v[t];
    if (tid6 == 0) mbar_init(macdone, WARPS_PER_GATE);
    // ---- prologue: gate pre-combination (tfhe.rs:27-71), rounding of (b, a) (tfhe.rs:97,107-108), acc_0 ----
    int nsteps = a.nsteps;
    if (EXTPROD) {
        nsteps = 1;
        const uint32_t* src = a.trlwe_in + (size_t)gate * 2048;
        const uint32_t* sub = a.trlwe_in0 ? a.trlwe_in0 + (size_t)gate * 2048 : nullptr;   // cmux: rep_1 - rep_0 (trgsw.rs:315-322)
        for (int k = tid6; k < 2048; k += THREADS_PER_GATE) acc[k] = sub ? src[k] - sub[k] : src[k];
    } else {
        uint32_t* lin = dh;
        const bool second = gate >= a.split;
        const long gsrc = second ? gate - a.split : gate;
        const uint32_t* q0 = second ? a.in0b : a.in0;
        const uint32_t* q1 = second ? a.in1b : a.in1;
        const uint32_t k0 = (uint32_t)(second ? a.c0b : a.c0), k1 = (uint32_t)(second ? a.c1b : a.c1), kb = second ? a.cbb : a.cb;
        const uint32_t* p0 = q0 + (size_t)gsrc * (LWE_N + 1);
        const uint32_t* p1 = q1 ? q1 + (size_t)gsrc * (LWE_N + 1) : nullptr;
        for (int c = tid6; c <= LWE_N; c += THREADS_PER_GATE) {
            uint32_t v = k0 * p0[c];
            if (p1) v += k1 * p1[c];
            if (c == 0) v += kb;
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
    if (!active) return;   // gate slots without a gate leave here: every barrier below is private to one gate
    // ---- 635 x CMUX ----
    // Synchronisation per step (named barriers, so gates sharing a CTA and the two polynomials of a gate decouple):
    //   poly barrier (96 threads)      : the masked source polynomial u[poly] (shared by its three digit warps) is complete
    //   B1 gate barrier (192 threads)  : the 6 digit spectra of this step are complete
    //   macdone (mbarrier, 6 arrivals) : every warp finished READING the digit spectra dh[] -> a warp may reuse its own
    //                                    plane dh[w6] as the transpose scratch of its inverse transform
    //   poly barrier (96 threads)      : the three key-slice warps of a polynomial have added their exact slices into
    //                                    acc[poly] (red.shared, no output planes: 24 KB less shared memory per gate)
    const int bar_gate = 1 + gl, bar_poly = 1 + G + 2 * gl + pw;
    uint32_t mac_parity = 0;
#pragma unroll 1
    for (int i = 0; i < nsteps; i++) {
        const uint32_t* step_bk = a.bkdev + (EXTPROD ? (size_t)(gate % a.ntrgsw) : (size_t)i) * bk_step_words(NS);
        uint32_t* S = dh + w6 * TILE_WORDS;
        {   // phase 1: a third of the rows of u[pw], then digit kw of u[pw] -> spectrum plane dh[w6]
            p1u<!EXTPROD>(lane, acc + pw * 1024, EXTPROD ? 0u : (uint32_t)abar[i], a.mask, kw, U + pw * 1024);
            bar_sync(bar_poly, 96);
            p1a(lane, U + pw * 1024, kw, S, dtab);
            __syncwarp();
            p1b(lane, S, twF);
        }
        bar_sync(bar_gate, THREADS_PER_GATE);
        uint32_t x[32];
        if (kw < NS) {   // phase 2: key slice kw of output poly pw (with two key slices the third warp of a polynomial only signals)
            p2a_mac_head(lane, step_bk + (size_t)(pw * NS + kw) * BK_SLAB_WORDS, dh, dh + 3 * TILE_WORDS, twI, x);
            if (EXTPROD) {   // plain external product: the result replaces acc; every warp clears its share before it arrives
#pragma unroll
                for (int r = kw; r < 32; r += 3) acc[pw * 1024 + 32 * r + lane] = 0u;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(macdone);
            gs32_tail(x, TwRow{twI + lane * TWB_STRIDE});
            gs_norm<2>(x);
            mbar_wait(macdone, mac_parity);
#pragma unroll
            for (int q = 0; q < 8; q++)
                *reinterpret_cast<uint4*>(S + swz_chunk(lane, q)) = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
            __syncwarp();
            p2b(lane, S, kw, x, NS);   // x[r] = exact slice value (already shifted) of coefficient 32 r + lane
            // phase 3: acc[pw] += x by shared-memory reductions (the slice warps of a polynomial add concurrently)
            const uint32_t A = smem_u32(acc + pw * 1024 + lane);
#pragma unroll
            for (int r = 0; r < 32; r++) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(A + 128u * r), "r"(x[r]) : "memory");
        } else {
            if (EXTPROD) {
#pragma unroll
                for (int r = kw; r < 32; r += 3) acc[pw * 1024 + 32 * r + lane] = 0u;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(macdone);
        }
        mac_parity ^= 1u;
        bar_sync(bar_poly, 96);   // acc[pw] is complete before the next step decomposes it
    }
    bar_sync(bar_gate, THREADS_PER_GATE);

    // ---- epilogue: sample_extract_index(0) (trlwe.rs:110-121) + key-switch digits (tlwe.rs:47-64) ----
    if (a.trlwe_out) {
        uint32_t* dst = a.trlwe_out + (size_t)gate * 2048;
        const uint32_t* add = (EXTPROD && a.trlwe_in0) ? a.trlwe_in0 + (size_t)gate * 2048 : nullptr;   // cmux: ... + rep_0
        for (int k = tid6; k < 2048; k += THREADS_PER_GATE) dst[k] = add ? acc[k] + add[k] : acc[k];
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
// K5L: latency shape of the blind rotation -- ONE gate on a cluster of TWO CTAs (two SMs), three warps each.
// A warp instruction stream of one CMUX step needs ~3900 FMA-pipe cycles of its SM sub-partition; with six warps on one
// SM two sub-partitions carry two warps and set the pace (measured 11.8 k cycles per step).  Here CTA `pw` of the pair
// owns polynomial pw (0 = b, 1 = a): its three warps sit on three different sub-partitions, its accumulator and the
// masked difference stay local, and the only exchange per step is the 12 KB of digit spectra, which every CTA also
// stores into its peer's shared memory (distributed shared memory, double buffered) before ONE cluster barrier.
// =====================================================================================================
namespace cg = cooperative_groups;
constexpr int PAIR_THREADS = 96;
constexpr int PAIR_SMEM_WORDS = TW_SMEM_WORDS + 1024 /*acc*/ + 1024 /*U*/ + 3 * TILE_WORDS /*own spectra*/ + 2 * 3 * TILE_WORDS /*peer spectra x2*/ + 320 +
                                2 * 3 * (int)BK_SLAB_WORDS /*key slabs of this and the next step, one per warp*/;
template <int NS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1) blind_rotate_pair_kernel(const BrArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int pw = (int)cluster.block_rank();
    const long gate = blockIdx.x >> 1;
    const int kw = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    uint32_t* twF = smem;
    uint32_t* twI = smem + 32 * TWB_STRIDE;
    uint32_t* acc = smem + TW_SMEM_WORDS;
    uint32_t* U = acc + 1024;
    uint32_t* own = U + 1024;            // [3] tiles: spectra of this CTA's polynomial; plane kw doubles as transpose scratch
    uint32_t* peer = own + 3 * TILE_WORDS;     // [2][3] tiles: spectra of the other polynomial, written by the other CTA
    uint16_t* abar = reinterpret_cast<uint16_t*>(peer + 6 * TILE_WORDS);
    uint64_t* macdone = reinterpret_cast<uint64_t*>(peer + 6 * TILE_WORDS + 318);
    uint32_t* slabs = peer + 6 * TILE_WORDS + 320;              // [2][3][BK_SLAB_WORDS]: warp-private, filled one step ahead by cp.async
    uint32_t* remote = cluster.map_shared_rank(peer, pw ^ 1);   // where MY spectra go in the other CTA
    // the key slab of step i for this warp: 48 x 512 B, copied asynchronously a whole step ahead so that no L2 round trip is
    // left on the critical path of a lone warp
    auto slab_fetch = [&](int step) {
        if (kw >= NS) { asm volatile("cp.async.commit_group;" ::: "memory"); return; }
        const uint4* src = reinterpret_cast<const uint4*>(a.bkdev + (size_t)step * bk_step_words(NS) + (size_t)(pw * NS + kw) * BK_SLAB_WORDS) + lane;
        const uint32_t dst = smem_u32(slabs + ((step & 1) * 3 + kw) * BK_SLAB_WORDS) + 16u * lane;
#pragma unroll
        for (int t = 0; t < 48; t++) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 512u * t), "l"(src + 32 * t) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    uint32_t* dtab = smem + 2 * 32 * TWB_STRIDE;
    for (int t = tid; t < 32 * TWB_STRIDE; t += PAIR_THREADS) { twF[t] = g_fwdB[t]; twI[t] = g_invB[t]; }
    if (tid < DIGIT_TAB_WORDS) dtab[tid] = g_digit_tab.v[tid];
    if (tid == 0) mbar_init(macdone, 3);
    {   // prologue (both CTAs read the inputs): gate pre-combination, rounding, acc_0 of the own polynomial
        uint32_t* lin = own;
        const bool second = gate >= a.split;
        const long gsrc = second ? gate - a.split : gate;
        const uint32_t* q0 = second ? a.in0b : a.in0;
        const uint32_t* q1 = second ? a.in1b : a.in1;
        const uint32_t k0 = (uint32_t)(second ? a.c0b : a.c0), k1 = (uint32_t)(second ? a.c1b : a.c1), kb = second ? a.cbb : a.cb;
        const uint32_t* p0 = q0 + (size_t)gsrc * (LWE_N + 1);
        const uint32_t* p1 = q1 ? q1 + (size_t)gsrc * (LWE_N + 1) : nullptr;
        for (int c = tid; c <= LWE_N; c += PAIR_THREADS) {
            uint32_t v = k0 * p0[c];
            if (p1) v += k1 * p1[c];
            if (c == 0) v += kb;
            lin[c] = v;
        }
        __syncthreads();
        for (int i = tid; i < LWE_N; i += PAIR_THREADS) abar[i] = (uint16_t)((lin[1 + i] + (1u << 20)) >> 21);
        const uint32_t bbar = lin[0] >> 21;
        const uint32_t nrot = (2048u - bbar) & 2047u;
        for (int k = tid; k < 1024; k += PAIR_THREADS) {
            const bool neg = ((uint32_t)k < (nrot & 1023u)) != (nrot >= 1024u);
            acc[k] = pw == 0 ? (neg ? 0u - a.mu : a.mu) : 0u;
        }
    }
    cluster.sync();   // both CTAs are set up (mbarriers, tables) before the first remote store
    if (a.nsteps > 0) slab_fetch(0);
    uint32_t mac_parity = 0;
#pragma unroll 1
    for (int i = 0; i < a.nsteps; i++) {
        uint32_t* S = own + kw * TILE_WORDS;
        uint32_t x[32];
        p1u<true>(lane, acc, (uint32_t)abar[i], a.mask, kw, U);
        bar_sync(1, PAIR_THREADS);
        p1a(lane, U, kw, S, dtab);
        __syncwarp();
        fwd_rows(lane, S, twF, x);
        {
            uint32_t* R = remote + ((i & 1) * 3 + kw) * TILE_WORDS;
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint4 v = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
                *reinterpret_cast<uint4*>(S + swz_chunk(lane, q)) = v;
                *reinterpret_cast<uint4*>(R + swz_chunk(lane, q)) = v;
            }
        }
        // split cluster barrier: arrive (release: my remote stores), start the copy of the NEXT step's key slab, then wait
        // (acquire).  (A point-to-point handshake on cluster-scope mbarriers was measured 10 % slower than this barrier.)
        cluster.barrier_arrive();
        if (i + 1 < a.nsteps) slab_fetch(i + 1); else asm volatile("cp.async.commit_group;" ::: "memory");
        {
            asm volatile("cp.async.wait_group 1;" ::: "memory");   // this step's slab (committed one step ago) has landed
            __syncwarp();
            const uint32_t* slab = slabs + ((i & 1) * 3 + kw) * BK_SLAB_WORDS;
            const uint32_t* P = peer + (i & 1) * 3 * TILE_WORDS;
            bar_sync(2, PAIR_THREADS);   // this CTA's own three spectra are complete (local barrier; the cluster one is still pending)
            if (kw < NS) {
                // the key rows that meet this CTA's own spectra need nothing from the peer: accumulate them while the cluster
                // barrier is pending, then wait and add the rows of the peer's spectra
                uint64_t mac[32];
                p2a_mac_part(lane, slab, own, pw == 0 ? 0 : 3, mac, true);
                cluster.barrier_wait();   // the peer's spectra are here; the peer has finished the previous step's MAC
                p2a_mac_part(lane, slab, P, pw == 0 ? 3 : 0, mac, false);
                p2a_mac_finish(lane, mac, twI, x);
                __syncwarp();
                if (lane == 0) mbar_arrive(macdone);
                gs32_tail(x, TwRow{twI + lane * TWB_STRIDE});
                gs_norm<2>(x);
                mbar_wait(macdone, mac_parity);
#pragma unroll
                for (int q = 0; q < 8; q++)
                    *reinterpret_cast<uint4*>(S + swz_chunk(lane, q)) = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
                __syncwarp();
                p2b(lane, S, kw, x, NS);
                const uint32_t A = smem_u32(acc + lane);
#pragma unroll
                for (int r = 0; r < 32; r++) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(A + 128u * r), "r"(x[r]) : "memory");
            } else {
                cluster.barrier_wait();
                if (lane == 0) mbar_arrive(macdone);
            }
            mac_parity ^= 1u;
        }
        bar_sync(1, PAIR_THREADS);
    }
    // ---- epilogue: CTA 0 owns b, CTA 1 owns a ----
    if (a.trlwe_out) {
        uint32_t* dst = a.trlwe_out + (size_t)gate * 2048 + pw * 1024;
        for (int k = tid; k < 1024; k += PAIR_THREADS) dst[k] = acc[k];
    }
    if (pw == 1 && (a.ksdig || a.lwe1_out)) {
        for (int i = tid; i < 1024; i += PAIR_THREADS) {
            const uint32_t ai = (i == 0) ? acc[0] : 0u - acc[1024 - i];
            if (a.ksdig) a.ksdig[(size_t)gate * 1024 + i] = (uint16_t)((ai + 0x8000u) >> 16);
            if (a.lwe1_out) a.lwe1_out[(size_t)gate * 1025 + 1 + i] = ai;
        }
    }
    if (pw == 0) {
        if (a.lwe1_out && tid == 0) a.lwe1_out[(size_t)gate * 1025] = acc[0];
        if (a.out_init) {
            uint32_t* dst = a.out_init + (size_t)gate * (LWE_N + 1);
            for (int c = tid; c <= LWE_N; c += PAIR_THREADS) dst[c] = (c == 0) ? acc[0] : 0u;
        }
    }
    cluster.sync();   // no CTA leaves while its peer may still store into its shared memory
}

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
                                                                         uint32_t* __restrict__ out, long B) {
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
        uint32_t* o = out + (size_t)(g0 + g) * (LWE_N + 1) + 4 * t;
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
// index = 30 KB) are copied once per CTA with cp.async into a 3-deep ring (one __syncthreads per stage) and consumed by
// the 16 gates of the tile.
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
constexpr int KS2_ROW_WORDS = 640;                 // 636 words padded to a multiple of 16 bytes x 32 lanes x 5
constexpr int KS2_RING = 3;
constexpr int KS2_STAGE_WORDS = KS2_ROWS * KS2_ROW_WORDS;
constexpr size_t KS2_SMEM_BYTES = (size_t)KS2_RING * KS2_STAGE_WORDS * 4 + (size_t)KS_ICHUNK * KS2_GATES * 2;
__global__ void __launch_bounds__(KS2_THREADS) keyswitch2_kernel(const uint4* __restrict__ ksk, const uint16_t* __restrict__ dig,
                                                                uint32_t* __restrict__ out, long B) {
    extern __shared__ __align__(16) uint32_t ks_smem[];
    uint32_t* ring = ks_smem;
    uint16_t* dg = reinterpret_cast<uint16_t*>(ks_smem + KS2_RING * KS2_STAGE_WORDS);   // [ichunk][16]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long g0 = (long)blockIdx.x * KS2_GATES;
    const int ichunk = 1024 / (int)gridDim.y;
    const int i0 = blockIdx.y * ichunk;
    const int nstages = ichunk * (8 / KS2_LV);
    for (int t = threadIdx.x; t < ichunk * KS2_GATES; t += KS2_THREADS) {
        const int g = t / ichunk, ii = t % ichunk;
        dg[ii * KS2_GATES + g] = (g0 + g < B) ? dig[(size_t)(g0 + g) * 1024 + i0 + ii] : (uint16_t)0;
    }
    auto stage_in = [&](int k) {   // rows (key index i0 + k / SPI, KS2_LV levels, all three multiples) -> ring slot k % 3
        constexpr int SPI = 8 / KS2_LV;   // stages per key index
        const uint4* src = ksk + ((size_t)(i0 + k / SPI) * 8 + (size_t)(k % SPI) * KS2_LV) * 3 * KS_CHUNKS;
        const uint32_t dst = smem_u32(ring + (k % KS2_RING) * KS2_STAGE_WORDS);
        for (int t = threadIdx.x; t < KS2_ROWS * KS_CHUNKS; t += KS2_THREADS) {
            const int row = t / KS_CHUNKS, c = t - row * KS_CHUNKS;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (uint32_t)(row * KS2_ROW_WORDS + 4 * c) * 4u), "l"(src + t)
                         : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage_in(0);
    if (nstages > 1) stage_in(1); else asm volatile("cp.async.commit_group;" ::: "memory");
    uint4 acc[5];
#pragma unroll
    for (int q = 0; q < 5; q++) acc[q] = make_uint4(0, 0, 0, 0);
    const bool live = g0 + warp < B;
#pragma unroll 1
    for (int k = 0; k < nstages; k++) {
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // stage k has landed (at most the newest group is still in flight)
        __syncthreads();                                          // ... for every thread; and everybody is done with stage k-1
        if (k + 2 < nstages) stage_in(k + 2); else asm volatile("cp.async.commit_group;" ::: "memory");
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
    uint32_t* o = out + (size_t)(g0 + warp) * (LWE_N + 1);
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
__global__ void bk_fill_kernel(uint32_t* __restrict__ bk, uint64_t seed, long nwords /* = rows * N */) {
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
__global__ void lwe_rows_kernel(uint32_t* __restrict__ out, long rows, uint64_t seed, uint64_t index0, const uint8_t* __restrict__ s0,
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

// =====================================================================================================
// host side: context + C ABI
// =====================================================================================================
// Work slots: every gate-batch call borrows one slot (key-switch digit workspace, staging buffers, scratch) from a
// small ring.  A slot is guarded by an event recorded behind the last operation that touches its buffers, so calls
// issued on DIFFERENT streams (or back-to-back asynchronous host calls) may overlap on the device: the tail of one
// batch's blind rotation and its key switch run under the head of the next batch.
struct Slot {
    cudaStream_t stream = nullptr;   // internal stream, used by the host-pointer entry points
    cudaEvent_t done = nullptr;
    bool pending = false;
    uint16_t* ksdig = nullptr; size_t ksdig_cap = 0;
    uint32_t* tmp[4] = {nullptr, nullptr, nullptr, nullptr}; size_t tmp_cap[4] = {0, 0, 0, 0};
    uint32_t* scratch = nullptr; size_t scratch_cap = 0;   // hom_mux intermediates / transformed TRGSWs of step-level calls
    uint8_t* s0buf = nullptr;                               // [1024] device copy of a caller's lv0 secret key (encrypt / decrypt)
};

static const size_t BK_TORUS_BYTES = (size_t)LWE_N * 12 * 1024 * 4;
static const size_t KSK_BYTES = (size_t)1024 * 8 * 3 * (LWE_N + 1) * 4;
struct tfhe_b200_ctx {
    tfhe_b200_params prm;
    int device = 0;
    int sm_count = 0;
    uint32_t* bkdev = nullptr;   // n * BK_STEP_WORDS
    uint32_t* kskdev = nullptr;  // [N][t][3][n+1]
    uint32_t* bk_torus = nullptr;  // [n][2l][2][N] torus-domain key as loaded / generated (kept for export: 31 MB)
    uint8_t* keybits = nullptr;    // device copy of (s0[n] | pad to 1024 | s1[N]) during device keygen
    int32_t* s1poly = nullptr;     // s1 as a polynomial of small integers, the multiplier of the a*s products
    bool have_bk = false, have_ksk = false;
    static constexpr int NSLOT = 4;
    Slot slots[NSLOT];
    unsigned next_slot = 0;
    static constexpr int RING = 64;          // event ring: per-launch device times of the last RING timed gate batches
    cudaEvent_t ev[RING][4] = {};
    uint64_t timed = 0;
    uint64_t launches = 0;
    uint64_t last_batch = 0;
    int gates_per_cta = 1;
    int variant = 7;  // blind-rotate launch shape, see launch_blind_rotate
    int key_slices = 3;  // 3 = exact in the worst case (default); 2 = opt-in fast mode (tfhe_b200_set_key_slices)
    int ks_variant = 2;  // key-switch kernel: 2 = rows staged in shared memory, one warp per gate; 1 = register tiles
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
#define RC(call)                                                                                          \
    do {                                                                                                  \
        int rc_ = (call);                                                                                 \
        if (rc_) return rc_;                                                                              \
    } while (0)

static int fail(tfhe_b200_ctx* ctx, int code, const char* msg) {
    if (ctx) ctx->err = msg; else g_create_err = msg;
    return code;
}
static int grow(tfhe_b200_ctx* ctx, void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return TFHE_B200_OK;
    if (*p) CK(cudaFree(*p));   // cudaFree waits for the device: nothing in flight can still use the old buffer
    *p = nullptr; *cap = 0;
    CK(cudaMalloc(p, bytes));
    *cap = bytes;
    return TFHE_B200_OK;
}
// borrow the next slot of the ring for work that will be enqueued on `st` (nullptr = the slot's own stream)
static int slot_acquire(tfhe_b200_ctx* ctx, cudaStream_t* st, bool own_stream, Slot** out) {
    Slot& s = ctx->slots[ctx->next_slot++ % tfhe_b200_ctx::NSLOT];
    if (own_stream) *st = s.stream;
    if (s.pending) CK(cudaStreamWaitEvent(*st, s.done, 0));
    *out = &s;
    return TFHE_B200_OK;
}
static int slot_release(tfhe_b200_ctx* ctx, Slot* s, cudaStream_t st) {
    CK(cudaEventRecord(s->done, st));
    s->pending = true;
    return TFHE_B200_OK;
}

extern "C" {
static int transform_keys(tfhe_b200_ctx* ctx, const uint32_t* src_dev, uint32_t* dst_dev, int nsteps, cudaStream_t st);
}
template <class Kern>
static cudaError_t set_smem(Kern k, int G) { return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)br_smem_bytes(G)); }

extern "C" {

const char* tfhe_b200_version(void) { return "rustfhe_b200 0.2 (sm_100a, p=536856577, 3x11-bit key slices)"; }

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
    auto bail = [&](const char* what, cudaError_t ee) { g_create_err = std::string(what) + ": " + cudaGetErrorString(ee); tfhe_b200_ctx_destroy(ctx); return TFHE_B200_ERR_CUDA; };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    for (auto& s : ctx->slots) {
        if ((e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
        if ((e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    }
    for (auto& slot : ctx->ev) for (auto& ev : slot) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc(&ctx->bkdev, (size_t)LWE_N * BK_STEP_WORDS * 4)) != cudaSuccess) return bail("cudaMalloc(bk)", e);
    if ((e = cudaMalloc(&ctx->kskdev, (size_t)1024 * 8 * 3 * (LWE_N + 1) * 4)) != cudaSuccess) return bail("cudaMalloc(ksk)", e);
    if ((e = cudaMalloc(&ctx->bk_torus, BK_TORUS_BYTES)) != cudaSuccess) return bail("cudaMalloc(bk_torus)", e);
    if ((e = cudaMalloc(&ctx->keybits, 2048)) != cudaSuccess) return bail("cudaMalloc(keybits)", e);
    if ((e = cudaMalloc(&ctx->s1poly, 1024 * 4)) != cudaSuccess) return bail("cudaMalloc(s1poly)", e);
    for (auto& s : ctx->slots) if ((e = cudaMalloc(&s.s0buf, 1024)) != cudaSuccess) return bail("cudaMalloc(s0buf)", e);
    if ((e = set_smem(blind_rotate_kernel<1, false, 1>, 1)) != cudaSuccess) return bail("smem attr", e);
    if ((e = set_smem(blind_rotate_kernel<1, false, 3>, 1)) != cudaSuccess) return bail("smem attr", e);
    if ((e = set_smem(blind_rotate_kernel<4, false, 1>, 4)) != cudaSuccess) return bail("smem attr", e);
    if ((e = set_smem(blind_rotate_kernel<2, true, 1>, 2)) != cudaSuccess) return bail("smem attr", e);
    if ((e = set_smem(blind_rotate_kernel<4, true, 1>, 4)) != cudaSuccess) return bail("smem attr", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_pair_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_WORDS * 4)) != cudaSuccess)
        return bail("smem attr (pair)", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_WORDS * 4)) != cudaSuccess)
        return bail("smem attr (pair)", e);
    if ((e = set_smem(blind_rotate_kernel<4, false, 1, 2>, 4)) != cudaSuccess) return bail("smem attr", e);
    if ((e = set_smem(blind_rotate_kernel<1, false, 1, 2>, 1)) != cudaSuccess) return bail("smem attr", e);
    if ((e = set_smem(blind_rotate_kernel<2, true, 1, 2>, 2)) != cudaSuccess) return bail("smem attr", e);
    if ((e = set_smem(blind_rotate_kernel<4, true, 1, 2>, 4)) != cudaSuccess) return bail("smem attr", e);
    if (const char* v = getenv("TFHE_B200_BR_VARIANT")) ctx->variant = atoi(v);
    if (const char* v = getenv("TFHE_B200_KS_VARIANT")) ctx->ks_variant = atoi(v);
    if (const char* v = getenv("TFHE_B200_KEY_SLICES")) ctx->key_slices = (atoi(v) == 2) ? 2 : 3;
    if ((e = cudaFuncSetAttribute(keyswitch2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KS2_SMEM_BYTES)) != cudaSuccess)
        return bail("smem attr (keyswitch2)", e);
    *out = ctx;
    return TFHE_B200_OK;
}

int tfhe_b200_ctx_destroy(tfhe_b200_ctx* ctx) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(ctx->bkdev); cudaFree(ctx->kskdev); cudaFree(ctx->bk_torus); cudaFree(ctx->keybits); cudaFree(ctx->s1poly);
    for (auto& s : ctx->slots) {
        cudaFree(s.ksdig); cudaFree(s.scratch); cudaFree(s.s0buf);
        for (auto p : s.tmp) cudaFree(p);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    for (auto& slot : ctx->ev) for (auto ev : slot) if (ev) cudaEventDestroy(ev);
    delete ctx;
    return TFHE_B200_OK;
}

const char* tfhe_b200_last_error(const tfhe_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int tfhe_b200_set_decomp_mask(tfhe_b200_ctx* ctx, uint32_t mask) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    ctx->prm.decomp_mask = mask;
    return TFHE_B200_OK;
}

// Pre-allocate every work slot for batches of up to max_batch gates (hom_mux needs twice the key-switch workspace), so
// that no call in a latency- or throughput-critical region ever reaches cudaMalloc (which blocks while the GPU is busy).
int tfhe_b200_reserve(tfhe_b200_ctx* ctx, size_t max_batch) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    if (max_batch == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t ct = max_batch * (size_t)(LWE_N + 1) * 4;
    for (auto& s : ctx->slots) {
        RC(grow(ctx, (void**)&s.ksdig, &s.ksdig_cap, 2 * max_batch * 1024 * sizeof(uint16_t)));
        for (int k = 0; k < 4; k++) RC(grow(ctx, (void**)&s.tmp[k], &s.tmp_cap[k], ct));
        RC(grow(ctx, (void**)&s.scratch, &s.scratch_cap, 2 * ct));
    }
    return TFHE_B200_OK;
}

// Key slices per bootstrapping-key polynomial.  3 (default): 11/11/10-bit slices, every slice product is below p/2 in the WORST
// case, the result is the exact external product unconditionally.  2 (opt-in fast mode): 16/16-bit slices, two inverse
// transforms and a third of the pointwise work less per CMUX; a slice product then stays below p/2 only with overwhelming
// probability over the key's masks (9.8 sigma: about 1e-22 per coefficient, 3e-16 per gate), not in the worst case.
// Re-transforms the loaded bootstrapping key.
int tfhe_b200_set_key_slices(tfhe_b200_ctx* ctx, int slices) {
    if (!ctx || (slices != 2 && slices != 3)) return fail(ctx, TFHE_B200_ERR_PARAM, "set_key_slices: 2 or 3");
    if (slices == ctx->key_slices) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    ctx->key_slices = slices;
    if (ctx->have_bk) {
        cudaStream_t st = ctx->slots[0].stream;
        RC(transform_keys(ctx, ctx->bk_torus, ctx->bkdev, LWE_N, st));
        CK(cudaStreamSynchronize(st));
    }
    return TFHE_B200_OK;
}

int tfhe_b200_sync(tfhe_b200_ctx* ctx) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    CK(cudaSetDevice(ctx->device));
    for (auto& s : ctx->slots)
        if (s.pending) { CK(cudaEventSynchronize(s.done)); s.pending = false; }
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
    out->device_key_bytes = (uint64_t)LWE_N * bk_step_words(ctx->key_slices) * 4 + (uint64_t)1024 * 8 * 3 * (LWE_N + 1) * 4;
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
    bk_transform_kernel<<<(npolys + KT_WARPS - 1) / KT_WARPS, KT_WARPS * 32, 0, st>>>(src_dev, dst_dev, npolys, ctx->key_slices);
    ctx->launches++;
    CK(cudaGetLastError());
    return TFHE_B200_OK;
}
int tfhe_b200_load_bk_device(tfhe_b200_ctx* ctx, const uint32_t* bk_dev, void* stream) {
    if (!ctx || !bk_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "load_bk_device: null argument");
    CK(cudaSetDevice(ctx->device));
    if (bk_dev != ctx->bk_torus) CK(cudaMemcpyAsync(ctx->bk_torus, bk_dev, BK_TORUS_BYTES, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    RC(transform_keys(ctx, ctx->bk_torus, ctx->bkdev, LWE_N, (cudaStream_t)stream));
    ctx->have_bk = true;
    return TFHE_B200_OK;
}
int tfhe_b200_load_bk(tfhe_b200_ctx* ctx, const uint32_t* bk_host) {
    if (!ctx || !bk_host) return fail(ctx, TFHE_B200_ERR_PARAM, "load_bk: null argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    CK(cudaMemcpyAsync(ctx->bk_torus, bk_host, BK_TORUS_BYTES, cudaMemcpyHostToDevice, st));
    RC(tfhe_b200_load_bk_device(ctx, ctx->bk_torus, st));
    CK(cudaStreamSynchronize(st));
    return TFHE_B200_OK;
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
    cudaStream_t st = ctx->slots[0].stream;
    CK(cudaMemcpyAsync(ctx->kskdev, ksk_host, (size_t)1024 * 8 * 3 * (LWE_N + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    ctx->have_ksk = true;
    return TFHE_B200_OK;
}

}  // extern "C" (helpers below are internal)

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
    a.bkdev = ctx->bkdev; a.mask = ctx->prm.decomp_mask; a.mu = ctx->prm.mu; a.ns = ctx->key_slices;
    if (a.split <= 0) a.split = a.B;
    const int slot = (int)(ctx->timed % tfhe_b200_ctx::RING);
    if (timed) CK(cudaEventRecord(ctx->ev[slot][0], st));
    // Launch shapes.  B <= #SMs: one gate per CTA, 254 registers (latency shape).  Larger batches: CTAs of up to G gates
    // (default G = 4: 24 warps = 6 per SM sub-partition, 80 registers, one CTA per SM -- measured 57.6 k gates/s against
    // 50.7 k for three 1-gate CTAs per SM whose 18 warps load the four sub-partitions 5/5/4/4); the batch is dealt out
    // evenly over rounds * #SMs CTAs so that a batch that is not a multiple of G * #SMs ends with 3-gate CTAs instead of
    // a half-empty last wave.
    const bool full = a.B > (long)ctx->sm_count;
    const int variant = (a.ns == 2 && ctx->variant != 9) ? 7 : ctx->variant;   // the measured alternatives exist for three slices only
    auto deal = [&](int G) {
        const long cap = (long)G * ctx->sm_count;
        const long rounds = (a.B + cap - 1) / cap;
        long nctas = rounds * ctx->sm_count;
        if (nctas > a.B) nctas = a.B;
        a.cta_base = (int)(a.B / nctas); a.cta_rem = (int)(a.B % nctas);
        ctx->gates_per_cta = G;
        return (unsigned)nctas;
    };
    auto fixed = [&](int G) {   // every CTA owns exactly G gates (the last one possibly fewer)
        const long nctas = (a.B + G - 1) / G;
        a.cta_base = (int)(a.B / nctas); a.cta_rem = (int)(a.B % nctas);
        ctx->gates_per_cta = G;
        return (unsigned)nctas;
    };
    if (full && variant == 3) {   // the earlier default, kept selectable for A/B runs: three 1-gate CTAs per SM (96 registers)
        blind_rotate_kernel<1, false, 3><<<fixed(1), THREADS_PER_GATE, br_smem_bytes(1), st>>>(a);
    } else if (full) {   // default (variant 7)
        const unsigned grid = deal(4);
        if (a.ns == 2) blind_rotate_kernel<4, false, 1, 2><<<grid, 4 * THREADS_PER_GATE, br_smem_bytes(4), st>>>(a);
        else blind_rotate_kernel<4, false, 1, 3><<<grid, 4 * THREADS_PER_GATE, br_smem_bytes(4), st>>>(a);
    } else if (3 * a.B <= (long)ctx->sm_count && variant != 9) {   // latency shape: one gate on a cluster of two SMs (measured
                                                                   // better than one CTA per gate up to about #SMs/3 gates)
        a.cta_base = 1; a.cta_rem = 0;
        ctx->gates_per_cta = 1;
        if (a.ns == 2) blind_rotate_pair_kernel<2><<<(unsigned)(2 * a.B), PAIR_THREADS, (size_t)PAIR_SMEM_WORDS * 4, st>>>(a);
        else blind_rotate_pair_kernel<3><<<(unsigned)(2 * a.B), PAIR_THREADS, (size_t)PAIR_SMEM_WORDS * 4, st>>>(a);
    } else {
        if (a.ns == 2) blind_rotate_kernel<1, false, 1, 2><<<fixed(1), THREADS_PER_GATE, br_smem_bytes(1), st>>>(a);
        else blind_rotate_kernel<1, false, 1><<<fixed(1), THREADS_PER_GATE, br_smem_bytes(1), st>>>(a);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    if (timed) CK(cudaEventRecord(ctx->ev[slot][1], st));
    return TFHE_B200_OK;
}
static int launch_keyswitch(tfhe_b200_ctx* ctx, const uint16_t* dig, uint32_t* out, long B, cudaStream_t st, bool timed) {
    const int slot = (int)(ctx->timed % tfhe_b200_ctx::RING);
    if (timed) CK(cudaEventRecord(ctx->ev[slot][2], st));
    if (ctx->ks_variant == 1) {   // register-tile kernel (TFHE_B200_KS_VARIANT=1)
        const long tile = (long)KS_GT * KS_GROUPS;
        const long tiles = (B + tile - 1) / tile;
        int isplit = KS_ISPLIT_MIN;   // small batches (latency path, narrow circuit levels): split the key indices further to fill the SMs
        while (isplit < 128 && tiles * isplit < 2L * ctx->sm_count) isplit *= 2;
        dim3 grid((unsigned)tiles, isplit);
        keyswitch_kernel<<<grid, dim3(KS_THREADS, KS_GROUPS), 0, st>>>(reinterpret_cast<const uint4*>(ctx->kskdev), dig, out, B);
    } else {                      // shared-memory staged kernel, one warp per gate (default)
        const long tiles = (B + KS2_GATES - 1) / KS2_GATES;
        int isplit = KS_ISPLIT_MIN;
        while (isplit < 128 && tiles * isplit < 2L * ctx->sm_count) isplit *= 2;
        dim3 grid((unsigned)tiles, isplit);
        keyswitch2_kernel<<<grid, KS2_THREADS, KS2_SMEM_BYTES, st>>>(reinterpret_cast<const uint4*>(ctx->kskdev), dig, out, B);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    if (timed) { CK(cudaEventRecord(ctx->ev[slot][3], st)); ctx->timed++; }
    return TFHE_B200_OK;
}
static const size_t CT_WORDS = (size_t)(LWE_N + 1);
static const size_t CT_BYTES = CT_WORDS * 4;

// one bootstrapped gate batch on `st` with the workspaces of slot `s` (device pointers)
static int run_gates(tfhe_b200_ctx* ctx, Slot* s, BrArgs& a, uint32_t* out, cudaStream_t st) {
    RC(grow(ctx, (void**)&s->ksdig, &s->ksdig_cap, (size_t)a.B * 1024 * sizeof(uint16_t)));
    a.nsteps = LWE_N; a.out_init = out; a.ksdig = s->ksdig;
    RC(launch_blind_rotate(ctx, a, st, true));
    RC(launch_keyswitch(ctx, s->ksdig, out, a.B, st, true));
    ctx->last_batch = (uint64_t)a.B;
    return TFHE_B200_OK;
}
static int gate_args(tfhe_b200_ctx* ctx, const char* who, int op, const uint32_t* in0, const uint32_t* in1, size_t B, BrArgs* a) {
    if (op < 0 || op > TFHE_B200_ANDNY) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch: bad opcode");
    if (!ctx->have_bk || !ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "gate_batch: keys not loaded");
    bool two;
    op_coeffs(op, ctx->prm.mu, &a->c0, &a->c1, &a->cb, &two);
    if (two && !in1) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch: in1 required for this opcode");
    a->in0 = in0; a->in1 = two ? in1 : nullptr; a->B = (long)B;
    (void)who;
    return TFHE_B200_OK;
}
// hom_mux on device pointers: stage 1 = ONE launch of 2B gates (AND(control, in1) | AND(-control, in0)), stage 2 = OR
static int run_mux(tfhe_b200_ctx* ctx, Slot* s, const uint32_t* control, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B,
                   cudaStream_t st) {
    if (!ctx->have_bk || !ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "mux_batch: keys not loaded");
    RC(grow(ctx, (void**)&s->scratch, &s->scratch_cap, 2 * B * CT_BYTES));
    uint32_t* t = s->scratch;   // [0,B) = i_1 = hom_and(control, input_1); [B,2B) = i_0 = hom_and(-control, input_0)   (tfhe.rs:33-34)
    BrArgs a{};
    bool two;
    op_coeffs(TFHE_B200_AND, ctx->prm.mu, &a.c0, &a.c1, &a.cb, &two);
    op_coeffs(TFHE_B200_ANDNY, ctx->prm.mu, &a.c0b, &a.c1b, &a.cbb, &two);
    a.in0 = control; a.in1 = in1; a.in0b = control; a.in1b = in0; a.split = (long)B; a.B = (long)(2 * B);
    RC(run_gates(ctx, s, a, t, st));
    BrArgs b{};
    op_coeffs(TFHE_B200_OR, ctx->prm.mu, &b.c0, &b.c1, &b.cb, &two);   // bootstrap(i_1 + i_0 + 1/8)   (tfhe.rs:35-39)
    b.in0 = t; b.in1 = t + B * CT_WORDS; b.B = (long)B;
    return run_gates(ctx, s, b, out, st);
}

extern "C" {

int tfhe_b200_gate_batch_device(tfhe_b200_ctx* ctx, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B,
                                void* stream) {
    if (!ctx || !in0 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch: null argument");
    BrArgs a{};
    RC(gate_args(ctx, "gate_batch_device", op, in0, in1, B, &a));
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Slot* s;
    RC(slot_acquire(ctx, &st, false, &s));
    RC(run_gates(ctx, s, a, out, st));
    return slot_release(ctx, s, st);
}

// host pointers, asynchronous: H2D, gate batch, D2H are enqueued on one of the ctx's internal streams and the call
// returns; tfhe_b200_sync waits.  Host buffers should be pinned (pageable memory makes the copies synchronous).
int tfhe_b200_gate_batch_async(tfhe_b200_ctx* ctx, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B) {
    if (!ctx || !in0 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch: null argument");
    BrArgs a{};
    RC(gate_args(ctx, "gate_batch", op, in0, in1, B, &a));
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = nullptr;
    Slot* s;
    RC(slot_acquire(ctx, &st, true, &s));
    const size_t bytes = B * CT_BYTES;
    RC(grow(ctx, (void**)&s->tmp[0], &s->tmp_cap[0], bytes));
    RC(grow(ctx, (void**)&s->tmp[3], &s->tmp_cap[3], bytes));
    CK(cudaMemcpyAsync(s->tmp[0], in0, bytes, cudaMemcpyHostToDevice, st));
    a.in0 = s->tmp[0];
    if (a.in1) {
        RC(grow(ctx, (void**)&s->tmp[1], &s->tmp_cap[1], bytes));
        CK(cudaMemcpyAsync(s->tmp[1], in1, bytes, cudaMemcpyHostToDevice, st));
        a.in1 = s->tmp[1];
    }
    RC(run_gates(ctx, s, a, s->tmp[3], st));
    CK(cudaMemcpyAsync(out, s->tmp[3], bytes, cudaMemcpyDeviceToHost, st));
    return slot_release(ctx, s, st);
}
int tfhe_b200_gate_batch(tfhe_b200_ctx* ctx, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B) {
    RC(tfhe_b200_gate_batch_async(ctx, op, in0, in1, out, B));
    return tfhe_b200_sync(ctx);
}
int tfhe_b200_bootstrap_batch(tfhe_b200_ctx* ctx, const uint32_t* in, uint32_t* out, size_t B) {
    return tfhe_b200_gate_batch(ctx, TFHE_B200_COPY, in, nullptr, out, B);
}

int tfhe_b200_mux_batch_device(tfhe_b200_ctx* ctx, const uint32_t* control, const uint32_t* in0, const uint32_t* in1, uint32_t* out,
                               size_t B, void* stream) {
    if (!ctx || !control || !in0 || !in1 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "mux_batch: null argument");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Slot* s;
    RC(slot_acquire(ctx, &st, false, &s));
    RC(run_mux(ctx, s, control, in0, in1, out, B, st));
    return slot_release(ctx, s, st);
}
int tfhe_b200_mux_batch(tfhe_b200_ctx* ctx, const uint32_t* control, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B) {
    if (!ctx || !control || !in0 || !in1 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "mux_batch: null argument");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = nullptr;
    Slot* s;
    RC(slot_acquire(ctx, &st, true, &s));
    const size_t bytes = B * CT_BYTES;
    const uint32_t* src[3] = {control, in0, in1};
    for (int k = 0; k < 3; k++) {
        RC(grow(ctx, (void**)&s->tmp[k], &s->tmp_cap[k], bytes));
        CK(cudaMemcpyAsync(s->tmp[k], src[k], bytes, cudaMemcpyHostToDevice, st));
    }
    RC(grow(ctx, (void**)&s->tmp[3], &s->tmp_cap[3], bytes));
    RC(run_mux(ctx, s, s->tmp[0], s->tmp[1], s->tmp[2], s->tmp[3], B, st));
    CK(cudaMemcpyAsync(out, s->tmp[3], bytes, cudaMemcpyDeviceToHost, st));
    RC(slot_release(ctx, s, st));
    return tfhe_b200_sync(ctx);
}

}  // extern "C"

// ---- device-side key generation, encryption, decryption ----
extern "C" {

int tfhe_b200_keygen_device(tfhe_b200_ctx* ctx, uint64_t seed, const uint8_t* s0, const uint8_t* s1) {
    if (!ctx || !s0 || !s1) return fail(ctx, TFHE_B200_ERR_PARAM, "keygen_device: null argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    uint8_t hb[2048];
    int32_t hp[1024];
    memset(hb, 0, sizeof hb);
    memcpy(hb, s0, LWE_N);
    memcpy(hb + 1024, s1, 1024);
    for (int k = 0; k < 1024; k++) hp[k] = s1[k] ? 1 : 0;
    CK(cudaMemcpyAsync(ctx->keybits, hb, sizeof hb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->s1poly, hp, sizeof hp, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // hb / hp live on this stack frame
    const int rows = LWE_N * 6;
    // BK rows: A uniform, B = noise; B += A * s1 (exact negacyclic product); gadget term; NTT-domain transform
    bk_fill_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->bk_torus, seed, (long)rows * 1024);
    polymul_kernel<<<(rows + PM_WARPS - 1) / PM_WARPS, PM_WARPS * 32, 0, st>>>(ctx->bk_torus + 1024, ctx->s1poly, ctx->bk_torus, rows, 2048, 0,
                                                                                 2048, 1);
    bk_gadget_kernel<<<(rows + 255) / 256, 256, 0, st>>>(ctx->bk_torus, ctx->keybits, rows);
    ctx->launches += 3;
    CK(cudaGetLastError());
    RC(transform_keys(ctx, ctx->bk_torus, ctx->bkdev, LWE_N, st));
    // KSK rows straight into the device key
    const long krows = 1024L * 8 * 3;
    lwe_rows_kernel<<<(unsigned)((krows + 7) / 8), 256, 0, st>>>(ctx->kskdev, krows, seed, 0, ctx->keybits, ctx->keybits + 1024, nullptr, 0);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
    ctx->have_bk = ctx->have_ksk = true;
    return TFHE_B200_OK;
}
int tfhe_b200_export_bk(tfhe_b200_ctx* ctx, uint32_t* bk_host) {
    if (!ctx || !bk_host) return fail(ctx, TFHE_B200_ERR_PARAM, "export_bk: null argument");
    if (!ctx->have_bk) return fail(ctx, TFHE_B200_ERR_STATE, "export_bk: bootstrapping key not loaded");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(bk_host, ctx->bk_torus, BK_TORUS_BYTES, cudaMemcpyDeviceToHost));
    return TFHE_B200_OK;
}
// device-to-device forms: the source of the one-off NCCL broadcast that replicates the keys to the other GPUs
int tfhe_b200_export_bk_device(tfhe_b200_ctx* ctx, uint32_t* bk_dev, void* stream) {
    if (!ctx || !bk_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "export_bk_device: null argument");
    if (!ctx->have_bk) return fail(ctx, TFHE_B200_ERR_STATE, "export_bk_device: bootstrapping key not loaded");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(bk_dev, ctx->bk_torus, BK_TORUS_BYTES, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return TFHE_B200_OK;
}
int tfhe_b200_export_ksk_device(tfhe_b200_ctx* ctx, uint32_t* ksk_dev, void* stream) {
    if (!ctx || !ksk_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "export_ksk_device: null argument");
    if (!ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "export_ksk_device: key-switching key not loaded");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ksk_dev, ctx->kskdev, KSK_BYTES, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return TFHE_B200_OK;
}
int tfhe_b200_export_ksk(tfhe_b200_ctx* ctx, uint32_t* ksk_host) {
    if (!ctx || !ksk_host) return fail(ctx, TFHE_B200_ERR_PARAM, "export_ksk: null argument");
    if (!ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "export_ksk: key-switching key not loaded");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(ksk_host, ctx->kskdev, KSK_BYTES, cudaMemcpyDeviceToHost));
    return TFHE_B200_OK;
}
// bits_dev: [B] bytes on the device; out_dev: [B][n+1] on the device; s0: host
int tfhe_b200_encrypt_bits_device(tfhe_b200_ctx* ctx, uint64_t seed, uint64_t ct_index0, const uint8_t* s0, const uint8_t* bits_dev,
                                  size_t B, uint32_t* out_dev, void* stream) {
    if (!ctx || !s0 || (B && (!bits_dev || !out_dev))) return fail(ctx, TFHE_B200_ERR_PARAM, "encrypt_bits_device: null argument");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Slot* s;
    RC(slot_acquire(ctx, &st, false, &s));
    CK(cudaMemcpyAsync(s->s0buf, s0, LWE_N, cudaMemcpyHostToDevice, st));
    lwe_rows_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(out_dev, (long)B, seed, ct_index0, s->s0buf, nullptr, bits_dev, 1);
    ctx->launches++;
    CK(cudaGetLastError());
    return slot_release(ctx, s, st);
}
// phase_dev / bits_dev: either may be NULL
int tfhe_b200_decrypt_bits_device(tfhe_b200_ctx* ctx, const uint8_t* s0, const uint32_t* ct_dev, size_t B, uint8_t* bits_dev,
                                  uint32_t* phase_dev, void* stream) {
    if (!ctx || !s0 || (B && !ct_dev)) return fail(ctx, TFHE_B200_ERR_PARAM, "decrypt_bits_device: null argument");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Slot* s;
    RC(slot_acquire(ctx, &st, false, &s));
    CK(cudaMemcpyAsync(s->s0buf, s0, LWE_N, cudaMemcpyHostToDevice, st));
    lwe_phase_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(ct_dev, (long)B, s->s0buf, phase_dev, bits_dev);
    ctx->launches++;
    CK(cudaGetLastError());
    return slot_release(ctx, s, st);
}

}  // extern "C"

// external product / cmux on device pointers: transform the TRGSW samples into slot scratch, then one CMUX step per product
static int run_extprod(tfhe_b200_ctx* ctx, Slot* s, const uint32_t* trgsw_dev, size_t ntrgsw, const uint32_t* rep1, const uint32_t* rep0,
                       uint32_t* out, size_t B, cudaStream_t st) {
    RC(grow(ctx, (void**)&s->scratch, &s->scratch_cap, ntrgsw * BK_STEP_WORDS * 4));
    RC(transform_keys(ctx, trgsw_dev, s->scratch, (int)ntrgsw, st));
    BrArgs a{};
    a.bkdev = s->scratch; a.mask = ctx->prm.decomp_mask; a.mu = ctx->prm.mu; a.B = (long)B; a.split = (long)B; a.nsteps = 1;
    a.trlwe_in = rep1; a.trlwe_in0 = rep0; a.trlwe_out = out; a.ntrgsw = (long)ntrgsw; a.ns = ctx->key_slices;
    if (B > (size_t)ctx->sm_count) {   // throughput shape: 4 products per CTA
        const unsigned grid = (unsigned)((B + 3) / 4);
        a.cta_base = (int)(B / grid); a.cta_rem = (int)(B % grid);
        if (a.ns == 2) blind_rotate_kernel<4, true, 1, 2><<<grid, 4 * THREADS_PER_GATE, br_smem_bytes(4), st>>>(a);
        else blind_rotate_kernel<4, true, 1><<<grid, 4 * THREADS_PER_GATE, br_smem_bytes(4), st>>>(a);
    } else {
        const unsigned grid = (unsigned)((B + 1) / 2);
        a.cta_base = (int)(B / grid); a.cta_rem = (int)(B % grid);
        if (a.ns == 2) blind_rotate_kernel<2, true, 1, 2><<<grid, 2 * THREADS_PER_GATE, br_smem_bytes(2), st>>>(a);
        else blind_rotate_kernel<2, true, 1><<<grid, 2 * THREADS_PER_GATE, br_smem_bytes(2), st>>>(a);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return TFHE_B200_OK;
}

// ---- step-level entries: host pointers, synchronous.  `fn` enqueues the device work on (slot, stream). ----
struct HostIo {
    const void* in[3] = {nullptr, nullptr, nullptr};
    size_t in_bytes[3] = {0, 0, 0};
    void* out = nullptr;
    size_t out_bytes = 0;
};
template <class F>
static int with_host_io(tfhe_b200_ctx* ctx, const HostIo& io, F fn) {
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = nullptr;
    Slot* s;
    RC(slot_acquire(ctx, &st, true, &s));
    uint32_t* dev_in[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < 3; k++) {
        if (!io.in[k]) continue;
        RC(grow(ctx, (void**)&s->tmp[k], &s->tmp_cap[k], io.in_bytes[k]));
        CK(cudaMemcpyAsync(s->tmp[k], io.in[k], io.in_bytes[k], cudaMemcpyHostToDevice, st));
        dev_in[k] = s->tmp[k];
    }
    RC(grow(ctx, (void**)&s->tmp[3], &s->tmp_cap[3], io.out_bytes));
    RC(fn(s, st, dev_in, s->tmp[3]));
    CK(cudaMemcpyAsync(io.out, s->tmp[3], io.out_bytes, cudaMemcpyDeviceToHost, st));
    RC(slot_release(ctx, s, st));
    return tfhe_b200_sync(ctx);
}

extern "C" {

int tfhe_b200_blind_rotate_batch(tfhe_b200_ctx* ctx, const uint32_t* in, int nsteps, uint32_t* out_trlwe, size_t B) {
    if (!ctx || !in || !out_trlwe) return fail(ctx, TFHE_B200_ERR_PARAM, "blind_rotate_batch: null argument");
    if (nsteps < 0 || nsteps > LWE_N) return fail(ctx, TFHE_B200_ERR_PARAM, "blind_rotate_batch: nsteps out of range");
    if (!ctx->have_bk) return fail(ctx, TFHE_B200_ERR_STATE, "blind_rotate_batch: bootstrapping key not loaded");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = in; io.in_bytes[0] = B * CT_BYTES; io.out = out_trlwe; io.out_bytes = B * 2048 * 4;
    return with_host_io(ctx, io, [&](Slot*, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        BrArgs a{};
        a.c0 = 1; a.in0 = di[0]; a.nsteps = nsteps; a.B = (long)B; a.trlwe_out = dout;
        return launch_blind_rotate(ctx, a, st, false);
    });
}
int tfhe_b200_bootstrap_lv1_batch(tfhe_b200_ctx* ctx, const uint32_t* in, uint32_t* out_lwe1, size_t B) {
    if (!ctx || !in || !out_lwe1) return fail(ctx, TFHE_B200_ERR_PARAM, "bootstrap_lv1_batch: null argument");
    if (!ctx->have_bk) return fail(ctx, TFHE_B200_ERR_STATE, "bootstrap_lv1_batch: bootstrapping key not loaded");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = in; io.in_bytes[0] = B * CT_BYTES; io.out = out_lwe1; io.out_bytes = B * 1025 * 4;
    return with_host_io(ctx, io, [&](Slot*, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        BrArgs a{};
        a.c0 = 1; a.in0 = di[0]; a.nsteps = LWE_N; a.B = (long)B; a.lwe1_out = dout;
        return launch_blind_rotate(ctx, a, st, false);
    });
}
int tfhe_b200_keyswitch_batch(tfhe_b200_ctx* ctx, const uint32_t* lwe1, uint32_t* out, size_t B) {
    if (!ctx || !lwe1 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "keyswitch_batch: null argument");
    if (!ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "keyswitch_batch: key-switching key not loaded");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = lwe1; io.in_bytes[0] = B * 1025 * 4; io.out = out; io.out_bytes = B * CT_BYTES;
    return with_host_io(ctx, io, [&](Slot* s, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        RC(grow(ctx, (void**)&s->ksdig, &s->ksdig_cap, B * 1024 * sizeof(uint16_t)));
        lwe1_prepare_kernel<<<(unsigned)B, 256, 0, st>>>(di[0], s->ksdig, dout, (long)B);
        ctx->launches++;
        return launch_keyswitch(ctx, s->ksdig, dout, (long)B, st, false);
    });
}
int tfhe_b200_external_product_batch(tfhe_b200_ctx* ctx, const uint32_t* trgsw, size_t ntrgsw, const uint32_t* trlwe, uint32_t* out,
                                     size_t B) {
    if (!ctx || !trgsw || !trlwe || !out || ntrgsw == 0) return fail(ctx, TFHE_B200_ERR_PARAM, "external_product_batch: bad argument");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = trlwe; io.in_bytes[0] = B * 2048 * 4; io.in[1] = trgsw; io.in_bytes[1] = ntrgsw * 12 * 1024 * 4;
    io.out = out; io.out_bytes = B * 2048 * 4;
    return with_host_io(ctx, io, [&](Slot* s, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        return run_extprod(ctx, s, di[1], ntrgsw, di[0], nullptr, dout, B, st);
    });
}
// device-pointer forms of the two micro-benchmark entries (BASELINE config 3 measured device resident)
int tfhe_b200_external_product_batch_device(tfhe_b200_ctx* ctx, const uint32_t* trgsw_dev, size_t ntrgsw, const uint32_t* trlwe_dev,
                                            uint32_t* out_dev, size_t B, void* stream) {
    if (!ctx || !trgsw_dev || !trlwe_dev || !out_dev || ntrgsw == 0) return fail(ctx, TFHE_B200_ERR_PARAM, "external_product_batch_device: bad argument");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Slot* s;
    RC(slot_acquire(ctx, &st, false, &s));
    RC(run_extprod(ctx, s, trgsw_dev, ntrgsw, trlwe_dev, nullptr, out_dev, B, st));
    return slot_release(ctx, s, st);
}
int tfhe_b200_negacyclic_mul_batch_device(tfhe_b200_ctx* ctx, const uint32_t* a_dev, const int32_t* d_dev, uint32_t* out_dev, size_t B,
                                          void* stream) {
    if (!ctx || !a_dev || !d_dev || !out_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "negacyclic_mul_batch_device: null argument");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    polymul_kernel<<<(unsigned)((B + PM_WARPS - 1) / PM_WARPS), PM_WARPS * 32, 0, (cudaStream_t)stream>>>(a_dev, d_dev, out_dev, (long)B, 1024, 1024,
                                                                                                           1024, 0);
    ctx->launches++;
    CK(cudaGetLastError());
    return TFHE_B200_OK;
}
int tfhe_b200_negacyclic_mul_batch(tfhe_b200_ctx* ctx, const uint32_t* a, const int32_t* d, uint32_t* out, size_t B) {
    if (!ctx || !a || !d || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "negacyclic_mul_batch: null argument");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = a; io.in_bytes[0] = B * 4096; io.in[1] = d; io.in_bytes[1] = B * 4096; io.out = out; io.out_bytes = B * 4096;
    return with_host_io(ctx, io, [&](Slot*, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        polymul_kernel<<<(unsigned)((B + PM_WARPS - 1) / PM_WARPS), PM_WARPS * 32, 0, st>>>(di[0], (const int32_t*)di[1], dout, (long)B, 1024, 1024,
                                                                                             1024, 0);
        ctx->launches++;
        CK(cudaGetLastError());
        return TFHE_B200_OK;
    });
}

// TRGSWRep::cmux(rep_1, rep_0) = cross(rep_1 - rep_0) + rep_0  (trgsw.rs:315-322, 323-330); trgsw in the torus domain
int tfhe_b200_cmux_batch(tfhe_b200_ctx* ctx, const uint32_t* trgsw, size_t ntrgsw, const uint32_t* rep1, const uint32_t* rep0,
                         uint32_t* out, size_t B) {
    if (!ctx || !trgsw || !rep1 || !rep0 || !out || ntrgsw == 0) return fail(ctx, TFHE_B200_ERR_PARAM, "cmux_batch: bad argument");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = rep1; io.in_bytes[0] = B * 2048 * 4; io.in[1] = trgsw; io.in_bytes[1] = ntrgsw * 12 * 1024 * 4;
    io.in[2] = rep0; io.in_bytes[2] = B * 2048 * 4; io.out = out; io.out_bytes = B * 2048 * 4;
    return with_host_io(ctx, io, [&](Slot* s, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        return run_extprod(ctx, s, di[1], ntrgsw, di[0], di[2], dout, B, st);
    });
}
// TRLWERep::sample_extract_index(index) (trlwe.rs:110-121): [B][2][N] -> [B][N+1]
int tfhe_b200_sample_extract_batch(tfhe_b200_ctx* ctx, const uint32_t* trlwe, int index, uint32_t* out_lwe1, size_t B) {
    if (!ctx || !trlwe || !out_lwe1) return fail(ctx, TFHE_B200_ERR_PARAM, "sample_extract_batch: null argument");
    if (index < 0 || index >= 1024) return fail(ctx, TFHE_B200_ERR_PARAM, "sample_extract_batch: index out of range");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = trlwe; io.in_bytes[0] = B * 2048 * 4; io.out = out_lwe1; io.out_bytes = B * 1025 * 4;
    return with_host_io(ctx, io, [&](Slot*, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        sample_extract_kernel<<<(unsigned)B, 256, 0, st>>>(di[0], dout, (long)B, index);
        ctx->launches++;
        CK(cudaGetLastError());
        return TFHE_B200_OK;
    });
}

}  // extern "C"
