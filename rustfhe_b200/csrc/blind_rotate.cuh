// blind_rotate.cuh -- key transform and blind-rotation kernels of the B200 TFHE engine (included by engine.cu only).
//   K8  bk_transform_kernel       : torus-domain bootstrapping key -> NTT domain (replaces TRGSWRepF::from, trgsw.rs:68-76)
//   K5  blind_rotate_kernel       : gate pre-combination + 635 x CMUX + sample extract (tfhe.rs:27-113, trgsw.rs:264-322,
//                                   trlwe.rs:110-121), state resident in shared memory; also TRGSW (x) TRLWE / cmux (EXTPROD)
//   K5L blind_rotate_pair_kernel  : latency shape, one gate on a cluster of two CTAs
#pragma once
#include <cooperative_groups.h>
#include "../../include/tfhe_b200.h"
#include "cmux_steps.cuh"

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
    const int32_t* idx0;     // [B] or null: ROW INDICES into in0 / in1 / out_init instead of row = gate (device-resident circuit
    const int32_t* idx1;     //               evaluation: all three then point at the same wire table)
    const int32_t* idxo;
    const uint8_t* ops;      // [B] or null: per-gate opcode (a circuit level with mixed gates in ONE launch); overrides c0/c1/cb
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
__device__ __forceinline__ void bar_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void cluster_arrive_relaxed() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
// one elected lane: expect `bytes` on the mbarrier and start a bulk (TMA) copy global -> shared that completes on it
__device__ __forceinline__ void bulk_fetch(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {   // shared::cluster address of `saddr` in CTA `rank`
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// one elected lane: bulk copy shared (this CTA) -> shared of another CTA of the cluster, counting its bytes on THAT CTA's mbarrier
__device__ __forceinline__ void bulk_send(uint32_t dst_cluster, const void* src, uint32_t bytes, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
                 "r"(smem_u32(src)), "r"(bytes), "r"(bar_cluster)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}

// linear pre-combination of a gate opcode (tfhe.rs:27-71; same table as op_coeffs on the host): lin = k0*in0 + k1*in1 + (kb, 0)
__device__ __forceinline__ void gate_coeffs(int op, uint32_t mu, uint32_t& k0, uint32_t& k1, uint32_t& kb) {
    switch (op) {
    case TFHE_B200_NAND: k0 = 0u - 1u; k1 = 0u - 1u; kb = mu; break;
    case TFHE_B200_AND: k0 = 1; k1 = 1; kb = 0u - mu; break;
    case TFHE_B200_OR: k0 = 1; k1 = 1; kb = mu; break;
    case TFHE_B200_XOR: k0 = 2; k1 = 2; kb = 2u * mu; break;
    case TFHE_B200_NOT: k0 = 0u - 1u; k1 = 0; kb = 0; break;
    case TFHE_B200_ANDNY: k0 = 0u - 1u; k1 = 1; kb = 0u - mu; break;
    default: k0 = 1; k1 = 0; kb = 0; break;   // COPY
    }
}

// SLAB_TMA (one gate per CTA only): every phase-2 warp keeps a private 24 KB buffer for its key slab; the slab of the NEXT
// step is requested with one bulk (TMA) copy as soon as the warp has consumed the current one, so that with few warps per
// sub-partition no L2 round trip is left inside the multiply-accumulate.
constexpr size_t br_smem_bytes_tma() { return br_smem_bytes(1) + (size_t)WARPS_PER_GATE * BK_SLAB_WORDS * 4 + 64; }
template <int G, bool EXTPROD, int MINB, int NS = 3, bool SLAB_TMA = false>
__global__ void __launch_bounds__(G* THREADS_PER_GATE, MINB) blind_rotate_kernel(const BrArgs a) {
    static_assert(!SLAB_TMA || (G == 1 && !EXTPROD), "the slab buffers fit beside one gate only");
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
    uint64_t* slabbar = reinterpret_cast<uint64_t*>(smem + TW_SMEM_WORDS + G * GATE_SMEM_WORDS) + w6;   // SLAB_TMA: [6] + slabs [6][BK_SLAB_WORDS]
    uint32_t* myslab = smem + TW_SMEM_WORDS + G * GATE_SMEM_WORDS + 16 + w6 * BK_SLAB_WORDS;
    if (SLAB_TMA && lane == 0) {
        mbar_init(slabbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
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
        uint32_t k0 = (uint32_t)(second ? a.c0b : a.c0), k1 = (uint32_t)(second ? a.c1b : a.c1), kb = second ? a.cbb : a.cb;
        if (a.ops) gate_coeffs(a.ops[gate], a.mu, k0, k1, kb);
        const uint32_t* p0 = q0 + (size_t)(a.idx0 ? (long)a.idx0[gate] : gsrc) * (LWE_N + 1);
        const uint32_t* p1 = (q1 && k1 != 0) ? q1 + (size_t)(a.idx1 ? (long)a.idx1[gate] : gsrc) * (LWE_N + 1) : nullptr;
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
    if (SLAB_TMA && kw < NS && lane == 0 && nsteps > 0)
        bulk_fetch(myslab, a.bkdev + (size_t)(pw * NS + kw) * BK_SLAB_WORDS, (uint32_t)BK_SLAB_WORDS * 4u, slabbar);
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
            if (SLAB_TMA) {
                mbar_wait(slabbar, (uint32_t)i & 1u);   // this step's slab (requested a step ago) has landed
                p2a_mac_head<true>(lane, myslab, dh, dh + 3 * TILE_WORDS, twI, x);
                __syncwarp();                           // every lane has consumed the slab: request the next one into the same buffer
                if (lane == 0 && i + 1 < nsteps)
                    bulk_fetch(myslab, step_bk + bk_step_words(NS) + (size_t)(pw * NS + kw) * BK_SLAB_WORDS, (uint32_t)BK_SLAB_WORDS * 4u, slabbar);
            } else {
                p2a_mac_head(lane, step_bk + (size_t)(pw * NS + kw) * BK_SLAB_WORDS, dh, dh + 3 * TILE_WORDS, twI, x);
            }
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
        uint32_t* dst = a.out_init + (size_t)(a.idxo ? (long)a.idxo[gate] : gate) * (LWE_N + 1);
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
constexpr int PAIR_LAUNCH_THREADS = PAIR_THREADS;
constexpr int PAIR_SMEM_WORDS = TW_SMEM_WORDS + 1024 /*acc*/ + 1024 /*U*/ + 2 * 3 * TILE_WORDS /*own spectra x2*/ + 2 * 3 * TILE_WORDS /*peer spectra x2*/ +
                                464 + 2 * 3 * (int)BK_SLAB_WORDS /*key slabs of this and the next step, one per warp*/;
static_assert((TW_SMEM_WORDS + 2048) % 4 == 0 && TILE_WORDS % 4 == 0 && (BK_SLAB_WORDS * 4) % 16 == 0, "bulk copies need 16-byte alignment");
static_assert(PAIR_SMEM_WORDS * 4 + 1024 <= 227 * 1024, "cluster kernel exceeds the shared memory of an SM");
template <int NS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_LAUNCH_THREADS, 1) blind_rotate_pair_kernel(const BrArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int pw = (int)cluster.block_rank();
    const long gate = blockIdx.x >> 1;
    const int kw = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
    uint32_t* twF = smem;
    uint32_t* twI = smem + 32 * TWB_STRIDE;
    uint32_t* acc = smem + TW_SMEM_WORDS;
    uint32_t* U = acc + 1024;
    uint32_t* own2 = U + 1024;           // [2][3] tiles: spectra of this CTA's polynomial, by step parity (the tile of the other
                                         // parity is the transpose scratch of the inverse transform)
    uint32_t* peer = own2 + 6 * TILE_WORDS;    // [2][3] tiles: spectra of the other polynomial, copied in by the other CTA
    uint16_t* abar = reinterpret_cast<uint16_t*>(peer + 6 * TILE_WORDS);
    uint64_t* xbar = reinterpret_cast<uint64_t*>(peer + 6 * TILE_WORDS + 332);      // [2]: the peer's spectra of a step have landed
    uint64_t* slabbar = reinterpret_cast<uint64_t*>(peer + 6 * TILE_WORDS + 320);   // [2][3]: one per slab buffer
    uint32_t* twcF = peer + 6 * TILE_WORDS + 336;               // [64] + [64]: the warp-uniform column twiddles of the forward and
    uint32_t* twcI = twcF + 64;                                 // inverse transforms in the layout of a twiddle row
    uint32_t* slabs = peer + 6 * TILE_WORDS + 464;              // [2][3][BK_SLAB_WORDS]: warp-private, filled one step ahead
    const uint32_t remote = map_to_cta(smem_u32(peer), (uint32_t)(pw ^ 1));    // where MY spectra go in the other CTA
    const uint32_t remote_bar = map_to_cta(smem_u32(xbar), (uint32_t)(pw ^ 1));
    // the key slab of step i for this warp (24 KB, contiguous): ONE bulk (TMA) copy issued by one lane a whole step ahead,
    // completing on the buffer's mbarrier, so that no L2 round trip is left on the critical path of a lone warp (48 cp.async
    // per lane were measured 4-5 % slower from 2 gates up: their issue alone took 4 % of a step)
    auto slab_fetch = [&](int step) {
        if (kw >= NS || lane != 0) return;
        bulk_fetch(slabs + ((step & 1) * 3 + kw) * BK_SLAB_WORDS, a.bkdev + (size_t)step * bk_step_words(NS) + (size_t)(pw * NS + kw) * BK_SLAB_WORDS,
                   (uint32_t)BK_SLAB_WORDS * 4u, slabbar + (step & 1) * 3 + kw);
    };
    auto slab_wait = [&](int step) {
        if (kw < NS) mbar_wait(slabbar + (step & 1) * 3 + kw, (uint32_t)(step >> 1) & 1u);
    };

    for (int t = tid; t < 32 * TWB_STRIDE; t += PAIR_LAUNCH_THREADS) { twF[t] = g_fwdB[t]; twI[t] = g_invB[t]; }
    if (tid < 64) { twcF[tid] = c_fwdA[tid]; twcI[tid] = c_invA[tid]; }
    if (tid == 0) {
        for (int k = 0; k < 6; k++) mbar_init(slabbar + k, 1);
        mbar_init(xbar, 1);
        mbar_init(xbar + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {   // prologue (both CTAs read the inputs): gate pre-combination, rounding, acc_0 of the own polynomial
        uint32_t* lin = own2;
        const bool second = gate >= a.split;
        const long gsrc = second ? gate - a.split : gate;
        const uint32_t* q0 = second ? a.in0b : a.in0;
        const uint32_t* q1 = second ? a.in1b : a.in1;
        uint32_t k0 = (uint32_t)(second ? a.c0b : a.c0), k1 = (uint32_t)(second ? a.c1b : a.c1), kb = second ? a.cbb : a.cb;
        if (a.ops) gate_coeffs(a.ops[gate], a.mu, k0, k1, kb);
        const uint32_t* p0 = q0 + (size_t)(a.idx0 ? (long)a.idx0[gate] : gsrc) * (LWE_N + 1);
        const uint32_t* p1 = (q1 && k1 != 0) ? q1 + (size_t)(a.idx1 ? (long)a.idx1[gate] : gsrc) * (LWE_N + 1) : nullptr;
        for (int c = tid; c <= LWE_N; c += PAIR_LAUNCH_THREADS) {
            uint32_t v = k0 * p0[c];
            if (p1) v += k1 * p1[c];
            if (c == 0) v += kb;
            lin[c] = v;
        }
        __syncthreads();
        for (int i = tid; i < LWE_N; i += PAIR_LAUNCH_THREADS) abar[i] = (uint16_t)((lin[1 + i] + (1u << 20)) >> 21);
        const uint32_t bbar = lin[0] >> 21;
        const uint32_t nrot = (2048u - bbar) & 2047u;
        for (int k = tid; k < 1024; k += PAIR_LAUNCH_THREADS) {
            const bool neg = ((uint32_t)k < (nrot & 1023u)) != (nrot >= 1024u);
            acc[k] = pw == 0 ? (neg ? 0u - a.mu : a.mu) : 0u;
        }
    }
    cluster.sync();   // both CTAs are set up (mbarriers, tables) before the first remote store
    if (a.nsteps > 0) slab_fetch(0);
    // Exchange of the digit spectra, one step: every working warp copies its finished 4.5 KB tile into the peer's shared
    // memory with ONE bulk copy (shared::cta -> shared::cluster) that counts its bytes on the PEER's mbarrier; the peer waits
    // on that mbarrier only.  No remote stores, no acknowledgements, no fence and no cluster barrier (SASS of
    // barrier.cluster.arrive.release: MEMBAR.ALL.GPU + ERRBAR, 18 % of a step when every warp paid it, 9 % of waiting left
    // when a fourth warp paid it).  Buffers are reused every other step; that is safe without a reverse signal: the peer
    // sends the spectra of step i only after it has received mine of step i-1, i.e. after my step i-2 is consumed.
#pragma unroll 1
    for (int i = 0; i < a.nsteps; i++) {
        uint32_t* own = own2 + (i & 1) * 3 * TILE_WORDS;
        uint32_t* S = own + kw * TILE_WORDS;
        uint32_t* T = own2 + ((i + 1) & 1) * 3 * TILE_WORDS + kw * TILE_WORDS;   // scratch: last step's tile, long since consumed
        uint32_t x[32];
        p1u<true>(lane, acc, (uint32_t)abar[i], a.mask, kw, U);
        bar_sync(1, PAIR_THREADS);
        if (tid == 0) mbar_expect_tx(xbar + (i & 1), 3u * TILE_WORDS * 4u);   // arm this step's arrival of the peer's three tiles
        fwd_shared(lane, U, kw, S, twcF, twF, x);   // both passes through one code body: the step fits the instruction cache
#pragma unroll
        for (int q = 0; q < 8; q++) *reinterpret_cast<uint4*>(S + swz_chunk(lane, q)) = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // my tile rows are visible to the copy engine
        __syncwarp();
        if (lane == 0) bulk_send(remote + (uint32_t)(((i & 1) * 3 + kw) * TILE_WORDS) * 4u, S, (uint32_t)TILE_WORDS * 4u, remote_bar + 8u * (i & 1));
        if (i + 1 < a.nsteps) slab_fetch(i + 1);
        slab_wait(i);   // this step's slab (requested one step ago) has landed
        {
            const uint32_t* slab = slabs + ((i & 1) * 3 + kw) * BK_SLAB_WORDS;
            const uint32_t* P = peer + (i & 1) * 3 * TILE_WORDS;
            bar_sync(2, PAIR_THREADS);   // this CTA's own three spectra are complete
            if (kw < NS) {
                // the key rows that meet this CTA's own spectra need nothing from the peer: accumulate them while the peer's
                // tiles are in flight, then wait and add the rows of the peer's spectra
                uint64_t mac[32];
                p2a_mac_part(lane, slab, own, pw == 0 ? 0 : 3, mac, true);
                mbar_wait(xbar + (i & 1), (uint32_t)(i >> 1) & 1u);
                p2a_mac_part(lane, slab, P, pw == 0 ? 3 : 0, mac, false);
                p2a_mac_redc(mac, x);
                inv_shared(lane, x, T, twI, twcI, kw, NS);
                const uint32_t A = smem_u32(acc + lane);
#pragma unroll
                for (int r = 0; r < 32; r++) asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(A + 128u * r), "r"(x[r]) : "memory");
            }
        }
        bar_sync(1, PAIR_THREADS);
    }
    // ---- epilogue: CTA 0 owns b, CTA 1 owns a ----
    if (a.trlwe_out) {
        uint32_t* dst = a.trlwe_out + (size_t)gate * 2048 + pw * 1024;
        for (int k = tid; k < 1024; k += PAIR_LAUNCH_THREADS) dst[k] = acc[k];
    }
    if (pw == 1 && (a.ksdig || a.lwe1_out)) {
        for (int i = tid; i < 1024; i += PAIR_LAUNCH_THREADS) {
            const uint32_t ai = (i == 0) ? acc[0] : 0u - acc[1024 - i];
            if (a.ksdig) a.ksdig[(size_t)gate * 1024 + i] = (uint16_t)((ai + 0x8000u) >> 16);
            if (a.lwe1_out) a.lwe1_out[(size_t)gate * 1025 + 1 + i] = ai;
        }
    }
    if (pw == 0) {
        if (a.lwe1_out && tid == 0) a.lwe1_out[(size_t)gate * 1025] = acc[0];
        if (a.out_init) {
            uint32_t* dst = a.out_init + (size_t)(a.idxo ? (long)a.idxo[gate] : gate) * (LWE_N + 1);
            for (int c = tid; c <= LWE_N; c += PAIR_LAUNCH_THREADS) dst[c] = (c == 0) ? acc[0] : 0u;
        }
    }
    cluster.sync();   // no CTA leaves while its peer may still store into its shared memory
}

