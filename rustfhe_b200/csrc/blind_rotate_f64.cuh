// blind_rotate_f64.cuh -- K5F, the throughput blind rotation of the FFT64 arithmetic mode (included by engine.cu only).
//   gate pre-combination + 635 x CMUX + sample extract (tfhe.rs:27-113, trgsw.rs:264-322, trlwe.rs:110-121)
// One gate = ONE warp; eight gates per CTA, one CTA per SM (256 threads, 255 registers: eight gates fill the register file and
// the shared memory of an SM; two CTAs of four gates measured 8 % slower, profiles/README.md).  Per CMUX a warp runs six
// forward transforms (one per gadget digit), multiplies every spectrum into BOTH output accumulators while it is still in
// registers (2 x 16 complex values per lane = 128 registers that live for the whole step), and runs two inverse transforms:
// no spectrum is ever written to shared memory, no barrier between warps.  The gates of a CTA only meet at the key ring:
//   bootstrapping key: [step][row j][output o] chunks of 8 KB, in the order every warp consumes them.  One elected thread streams
//   them with bulk (TMA) copies, 16 KB (one row, both output polynomials) at a time, into a two-slot ring in shared memory (full /
//   empty mbarriers: a slot is refilled by the last warp that leaves it); the eight gates read a chunk from there, so the key crosses L2 -> SM once per CTA instead of once
//   per gate ("reused across a batch of gates in shared memory").
// See fft64.cuh for the transform and DESIGN.md sections 2, 3 and 5 for the operation counts and measurements.
#pragma once
#include "blind_rotate.cuh"
#include "t2_steps.cuh"
#include "fft64.cuh"

#if !defined(F64_GATES_DEF)
#define F64_GATES_DEF 8
#define F64_RING_DEF 2
#endif
constexpr int F64_GATES = F64_GATES_DEF;        // gates (= warps) per CTA
constexpr int F64_CTAS_PER_SM = 8 / F64_GATES;
constexpr int F64_RING = F64_RING_DEF;          // key chunks resident per CTA
constexpr int F64_SLOT_ELEMS = 2 * (int)F64_CHUNK_ELEMS;                 // one ring slot = both output polynomials of a row: 16 KB
constexpr int F64_CHUNK_BYTES = (int)(F64_SLOT_ELEMS * sizeof(cd16));
constexpr int F64_GATE_SMEM_BYTES = 2 * 1024 * 4 /*acc*/ + 512 * 16 /*transpose scratch*/ + 6 * 32 * 16 /*digit planes*/ + 640 * 2 /*abar*/;
constexpr int F64_SHARED_BYTES = ((F64_TAB_ELEMS + F64_UNTW_ROWS * 32) * 16 + F64_RING * F64_CHUNK_BYTES + 2 * F64_RING * 8 + F64_RING * 4 + 15) / 16 * 16;
constexpr size_t f64_smem_bytes() { return (size_t)F64_SHARED_BYTES + (size_t)F64_GATES * F64_GATE_SMEM_BYTES; }
static_assert(F64_CTAS_PER_SM * (f64_smem_bytes() + 1024) <= 227 * 1024, "two CTAs must fit the shared memory of one SM");

__device__ __forceinline__ void f64_exchange(const cd (&send)[8], cd (&recv)[8]) {
#pragma unroll
    for (int m = 0; m < 8; m++) {
        recv[m].re = __shfl_xor_sync(0xffffffffu, send[m].re, 1);
        recv[m].im = __shfl_xor_sync(0xffffffffu, send[m].im, 1);
    }
}
// forward transform of the 16 values a lane holds (natural order j = 32 r + lane) -> spectrum (position p = 16 lane + register):
// pass A (stages 0..3 in registers), transpose, stage 4 on the load side (each lane reads both operands of its 16 butterflies
// and computes its output of each), pass B (stages 5..8 in registers).  No lane-pair exchange.
// PREFETCH: the per-lane twiddle rows are requested before pass A (36 registers for the length of pass A and the transpose;
// +3 % in the throughput kernel)
template <bool PREFETCH = true>
__device__ __forceinline__ void f64_forward(int lane, cd (&x)[16], cd16* S, const cd16* tb, cd (&y)[16]) {
    F64TwB tw;
    if (PREFETCH) f64_fwd_twB(lane, tb, tw);
    f64_fwd_passA(x);
    f64_t1_store(lane, x, S);
    __syncwarp();
    if (!PREFETCH) f64_fwd_twB(lane, tb, tw);
    f64_t1_load_cross(lane, S, tw.w[0], y);
    __syncwarp();   // the scratch is free for the next transform
    f64_fwd_passB(y, tw);
}

// inverse transform of one output spectrum (destroyed), rounded to the exact integers and added to the accumulator polynomial
// REPLACE: the result overwrites the accumulator polynomial (plain external product) instead of being added to it
template <bool REPLACE = false>
__device__ __forceinline__ void f64_inverse_acc(int lane, cd (&sp)[16], cd16* S, const cd16* ta, const cd16* ut, uint32_t* ao) {
    cd v[16];
    {
        f64_inv_low(sp);
        cd send[8], recv[8];
        f64_x_send(lane, sp, send);
        f64_exchange(send, recv);
        f64_inv_x_bfly(lane, sp, recv, v);
    }
    F64TwA tw;
    f64_inv_twA(lane, ta, tw);   // before the transpose: the compiler cannot hoist loads above a warp barrier
    f64_t2_store(lane, v, S);
    __syncwarp();
    f64_t2_load(lane, S, v);
    __syncwarp();
    f64_inv_passA(v, tw);
    uint32_t lo[16], hi[16];
    f64_untwist_round(lane, v, ut, lo, hi);
#pragma unroll
    for (int r = 0; r < 16; r++) {
        if (REPLACE) { ao[32 * r + lane] = lo[r]; ao[512 + 32 * r + lane] = hi[r]; }
        else {   // one shared-memory reduction per word instead of load + add + store (-128 instructions per gate and CMUX, -0.7 % time)
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_u32(ao + 32 * r + lane)), "r"(lo[r]) : "memory");
            asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(smem_u32(ao + 512 + 32 * r + lane)), "r"(hi[r]) : "memory");
        }
    }
}

struct F64Ring {
    cd16* slot;        // [F64_RING][512]
    uint64_t* full;    // [F64_RING]
    uint64_t* empty;   // [F64_RING]
    uint32_t* left;    // [F64_RING] warps that have handed the slot back in the current round
    const cd16* key;   // chunk 0 of this gate batch's first step
    long total;        // chunks the CTA consumes
    int active;        // warps of the CTA that own a gate
    long period;       // 0: chunk n of the stream is chunk n of the key; > 0: chunk n is chunk n % period (one TRGSW used over and over)
    uint32_t full32;   // shared-window address of full[0]; empty[] and left[] follow it (set by f64_ring_addr)
};
__device__ __forceinline__ void f64_ring_addr(F64Ring& rg) { rg.full32 = smem_u32(rg.full); }   // after full / empty / left are laid out back to back
// consumer side: wait for chunk n, run `use(slot)`, hand the slot back.  The LAST warp to leave a slot refills it with chunk
// n + F64_RING: no warp is the producer, so the warps of a CTA may drift up to a ring apart (they are started staggered so that
// the transposes and key reads of one gate fall under the arithmetic of another).
__device__ __forceinline__ void mbar_wait32(uint32_t b, uint32_t parity) {   // the barrier by its shared-window address
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT32_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cta.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra WAIT32_%=;\n\t}" ::"r"(b), "r"(parity) : "memory");
}
template <class F>
__device__ __forceinline__ void f64_with_chunk(const F64Ring& rg, long n, int lane, F&& use) {
    const int s = (int)(n & (F64_RING - 1));
    const uint32_t par = (uint32_t)((n / F64_RING) & 1);
    mbar_wait32(rg.full32 + 8u * (uint32_t)s, par);   // 32-bit shared addresses, converted once per kernel (a generic pointer costs a conversion per use)
    use(rg.slot + (size_t)s * F64_SLOT_ELEMS);
    __syncwarp();
    if (lane == 0) {
        asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(rg.full32 + 8u * (uint32_t)(F64_RING + s)) : "memory");   // empty[s]
        uint32_t before;
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(before) : "r"(rg.full32 + 16u * (uint32_t)F64_RING + 4u * (uint32_t)s) : "memory");   // left[s]
        if (before == (uint32_t)(rg.active - 1)) {
            rg.left[s] = 0;
            if (n + F64_RING < rg.total) {
                mbar_wait(rg.empty + s, par);   // every warp's reads of the slot are ordered before the copy that overwrites it
                const long c = rg.period ? (n + F64_RING) % rg.period : n + F64_RING;
                bulk_fetch(rg.slot + (size_t)s * F64_SLOT_ELEMS, rg.key + (size_t)c * F64_SLOT_ELEMS, F64_CHUNK_BYTES, rg.full + s);
            }
        }
    }
}

__global__ void __launch_bounds__(F64_GATES * 32, F64_CTAS_PER_SM) blind_rotate_f64_kernel(const BrArgs a, const cd16* __restrict__ key, int stagger_ns) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd16* tab = reinterpret_cast<cd16*>(smem_raw);
    const cd16* tb = tab;                                             // forward pass B / exchange twiddles
    const cd16* ta = tab + F64_FWDB_ROWS * 32;                        // inverse stages 5..8
    F64Ring rg;
    const cd16* ut = tab + F64_TAB_ELEMS;                             // untwist
    rg.slot = tab + F64_TAB_ELEMS + F64_UNTW_ROWS * 32;
    rg.full = reinterpret_cast<uint64_t*>(rg.slot + (size_t)F64_RING * F64_SLOT_ELEMS);
    rg.empty = rg.full + F64_RING;
    rg.left = reinterpret_cast<uint32_t*>(rg.empty + F64_RING);
    f64_ring_addr(rg);
    const int gl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* gbase = smem_raw + F64_SHARED_BYTES + (size_t)gl * F64_GATE_SMEM_BYTES;
    uint32_t* acc = reinterpret_cast<uint32_t*>(gbase);
    cd16* S = reinterpret_cast<cd16*>(gbase + 2 * 1024 * 4);
    uint4* D = reinterpret_cast<uint4*>(gbase + 2 * 1024 * 4 + 512 * 16);   // digit planes [digit][re / im][lane]
    uint16_t* abar = reinterpret_cast<uint16_t*>(gbase + 2 * 1024 * 4 + 512 * 16 + 6 * 32 * 16);

    // gates are dealt out evenly: the first cta_rem CTAs own cta_base+1 consecutive gates, the others cta_base (<= F64_GATES)
    const long cta = blockIdx.x;
    const long first = cta * a.cta_base + (cta < a.cta_rem ? cta : a.cta_rem);
    const int cnt = a.cta_base + (cta < a.cta_rem ? 1 : 0);
    const bool active = gl < cnt;
    const long gate = active ? first + gl : a.B - 1;
    const int nsteps = a.nsteps;
    rg.key = key;
    rg.total = (long)nsteps * 6;
    rg.active = cnt;
    rg.period = 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < F64_RING; s++) { mbar_init(rg.full + s, 1); mbar_init(rg.empty + s, cnt); rg.left[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const double* g0 = g_f64_fwdB; const double* g1 = g_f64_invA; const double* g2 = g_f64_untw;
        double* t = reinterpret_cast<double*>(tab);
        for (int k = threadIdx.x; k < F64_FWDB_ROWS * 64; k += blockDim.x) t[k] = g0[k];
        for (int k = threadIdx.x; k < F64_INVA_ROWS * 64; k += blockDim.x) t[F64_FWDB_ROWS * 64 + k] = g1[k];
        for (int k = threadIdx.x; k < F64_UNTW_ROWS * 64; k += blockDim.x) t[(F64_FWDB_ROWS + F64_INVA_ROWS) * 64 + k] = g2[k];
    }
    // ---- prologue: gate pre-combination (tfhe.rs:27-71), rounding of (b, a) (tfhe.rs:97,107-108), acc_0 ----
    {
        uint32_t* lin = reinterpret_cast<uint32_t*>(S);
        const bool second = gate >= a.split;
        const long gsrc = second ? gate - a.split : gate;
        const uint32_t* q0 = second ? a.in0b : a.in0;
        const uint32_t* q1 = second ? a.in1b : a.in1;
        uint32_t k0 = (uint32_t)(second ? a.c0b : a.c0), k1 = (uint32_t)(second ? a.c1b : a.c1), kb = second ? a.cbb : a.cb;
        if (a.ops) gate_coeffs(a.ops[gate], a.mu, k0, k1, kb);
        const uint32_t* p0 = q0 + (size_t)(a.idx0 ? (long)a.idx0[gate] : gsrc) * (LWE_N + 1);
        const uint32_t* p1 = (q1 && k1 != 0) ? q1 + (size_t)(a.idx1 ? (long)a.idx1[gate] : gsrc) * (LWE_N + 1) : nullptr;
        for (int c = lane; c <= LWE_N; c += 32) {
            uint32_t v = k0 * p0[c];
            if (p1) v += k1 * p1[c];
            if (c == 0) v += kb;
            lin[c] = v;
        }
        __syncwarp();
        for (int i = lane; i < LWE_N; i += 32) abar[i] = (uint16_t)((lin[1 + i] + (1u << 20)) >> 21);   // round
        const uint32_t bbar = lin[0] >> 21;                                                             // floor
        const uint32_t nrot = (2048u - bbar) & 2047u;   // acc_0 = X^{-bbar} * (mu, ..., mu ; 0)
        for (int k = lane; k < 1024; k += 32) {
            const bool neg = ((uint32_t)k < (nrot & 1023u)) != (nrot >= 1024u);
            acc[k] = neg ? 0u - a.mu : a.mu;
            acc[1024 + k] = 0;
        }
    }
    __syncthreads();   // tables, mbarriers
    if (threadIdx.x == 0)
        for (long n = 0; n < F64_RING && n < rg.total; n++)
            bulk_fetch(rg.slot + (size_t)n * F64_SLOT_ELEMS, rg.key + (size_t)n * F64_SLOT_ELEMS, F64_CHUNK_BYTES, rg.full + n);
    if (!active) return;   // gate slots without a gate leave here: the empty barriers count the active warps only
    if (stagger_ns > 0 && gl > 0) __nanosleep((unsigned)(gl * stagger_ns));

    // ---- 635 x CMUX ----
    long n = 0;
#pragma unroll 1
    for (int i = 0; i < nsteps; i++) {
        cd s0[16], s1[16];   // spectra of the two output polynomials
#pragma unroll
        for (int k = 0; k < 16; k++) { s0[k].re = 0.0; s0[k].im = 0.0; s1[k].re = 0.0; s1[k].im = 0.0; }
        const uint32_t ab = abar[i];
#pragma unroll 1
        for (int pw = 0; pw < 2; pw++) {
            {   // masked source words of polynomial pw, ((X^abar acc - acc) + mask) ^ mask: lane-private, parked as three byte planes
                uint32_t u[32];
                t2_u<true>(lane, acc + pw * 1024, ab, a.mask, u);
                u4 re, im;
                f64_pack_plane<0>(u, re, im);
                D[0 * 32 + lane] = make_uint4(re.x, re.y, re.z, re.w); D[1 * 32 + lane] = make_uint4(im.x, im.y, im.z, im.w);
                f64_pack_plane<1>(u, re, im);
                D[2 * 32 + lane] = make_uint4(re.x, re.y, re.z, re.w); D[3 * 32 + lane] = make_uint4(im.x, im.y, im.z, im.w);
                f64_pack_plane<2>(u, re, im);
                D[4 * 32 + lane] = make_uint4(re.x, re.y, re.z, re.w); D[5 * 32 + lane] = make_uint4(im.x, im.y, im.z, im.w);
            }
#pragma unroll 1
            for (int dw = 0; dw < 3; dw++) {
                cd x[16], y[16];
                {
                    const uint4 a4 = D[(2 * dw) * 32 + lane], b4 = D[(2 * dw + 1) * 32 + lane];
                    u4 re, im;
                    re.x = a4.x; re.y = a4.y; re.z = a4.z; re.w = a4.w; im.x = b4.x; im.y = b4.y; im.z = b4.z; im.w = b4.w;
                    f64_digits(re, im, x);
                }
                f64_forward(lane, x, S, tb, y);
                f64_with_chunk(rg, n, lane, [&](const cd16* k) {
                    f64_mac(lane, y, k, s0);
                    f64_mac(lane, y, k + F64_CHUNK_ELEMS, s1);
                });
                n++;
            }
        }
        f64_inverse_acc(lane, s0, S, ta, ut, acc);
        f64_inverse_acc(lane, s1, S, ta, ut, acc + 1024);
        __syncwarp();   // acc is complete before the next step's rotated reads (other lanes' words)
    }

    // ---- epilogue: sample_extract_index(0) (trlwe.rs:110-121) + key-switch digits (tlwe.rs:47-64) ----
    if (a.trlwe_out) {
        uint32_t* dst = a.trlwe_out + (size_t)gate * 2048;
        for (int k = lane; k < 2048; k += 32) dst[k] = acc[k];
    }
    if (a.ksdig || a.lwe1_out) {
        for (int i = lane; i < 1024; i += 32) {
            const uint32_t ai = (i == 0) ? acc[1024] : 0u - acc[1024 + 1024 - i];
            if (a.ksdig) a.ksdig[(size_t)gate * 1024 + i] = (uint16_t)((ai + 0x8000u) >> 16);
            if (a.lwe1_out) a.lwe1_out[(size_t)gate * 1025 + 1 + i] = ai;
        }
        if (a.lwe1_out && lane == 0) a.lwe1_out[(size_t)gate * 1025] = acc[0];
    }
    if (a.out_init) {
        uint32_t* dst = a.out_init + (size_t)(a.idxo ? (long)a.idxo[gate] : gate) * (LWE_N + 1);
        for (int c = lane; c <= LWE_N; c += 32) dst[c] = (c == 0) ? acc[0] : 0u;
    }
}

// K8F: key transform into the FFT64 layout.  One warp per (step i, row j, output poly o); 1/512 folded in.
constexpr int KTF_WARPS = 4;
__global__ void __launch_bounds__(KTF_WARPS * 32) bk_transform_f64_kernel(const uint32_t* __restrict__ bk, cd16* __restrict__ dev, int npolys) {
    __shared__ __align__(16) cd16 tb[F64_FWDB_ROWS * 32];
    __shared__ __align__(16) cd16 scratch[KTF_WARPS][512];
    {
        double* t = reinterpret_cast<double*>(tb);
        for (int k = threadIdx.x; k < F64_FWDB_ROWS * 64; k += blockDim.x) t[k] = g_f64_fwdB[k];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pid = blockIdx.x * KTF_WARPS + warp;
    if (pid >= npolys) return;
    cd x[16], y[16];
    f64_key_input(lane, bk + (size_t)pid * 1024, x);
    f64_forward(lane, x, scratch[warp], tb, y);
    cd16* dst = dev + (size_t)pid * F64_CHUNK_ELEMS;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        cd16 v; v.re = __dmul_rn(y[k].re, F64_KEY_SCALE); v.im = __dmul_rn(y[k].im, F64_KEY_SCALE);
        dst[k * 32 + lane] = v;
    }
}

// =====================================================================================================
// K5FX: plain external products / cmux with ONE shared TRGSW in the FFT64 arithmetic (trgsw.rs:264-322; BASELINE config 3):
//   out[g] = TRGSW (x) (in[g] - in0[g]) + in0[g]        (in0 = null: out[g] = TRGSW (x) in[g])
// Persistent CTAs of eight warps, one product per warp and round; the six 16 KB key chunks of the TRGSW go round the same ring
// as in K5F (chunk n of the stream = chunk n % 6 of the key), so the key crosses L2 -> SM once per eight products.
// =====================================================================================================
__global__ void __launch_bounds__(F64_GATES * 32, 1) external_product_f64_kernel(const BrArgs a, const cd16* __restrict__ key) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd16* tab = reinterpret_cast<cd16*>(smem_raw);
    const cd16* tb = tab;
    const cd16* ta = tab + F64_FWDB_ROWS * 32;
    F64Ring rg;
    const cd16* ut = tab + F64_TAB_ELEMS;
    rg.slot = tab + F64_TAB_ELEMS + F64_UNTW_ROWS * 32;
    rg.full = reinterpret_cast<uint64_t*>(rg.slot + (size_t)F64_RING * F64_SLOT_ELEMS);
    rg.empty = rg.full + F64_RING;
    rg.left = reinterpret_cast<uint32_t*>(rg.empty + F64_RING);
    f64_ring_addr(rg);
    const int gl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* gbase = smem_raw + F64_SHARED_BYTES + (size_t)gl * F64_GATE_SMEM_BYTES;
    uint32_t* acc = reinterpret_cast<uint32_t*>(gbase);
    cd16* S = reinterpret_cast<cd16*>(gbase + 2 * 1024 * 4);
    uint4* D = reinterpret_cast<uint4*>(gbase + 2 * 1024 * 4 + 512 * 16);
    const long ngroups = (a.B + F64_GATES - 1) / F64_GATES;
    const long mine = blockIdx.x < ngroups ? (ngroups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // rounds of this CTA
    rg.key = key;
    rg.total = mine * 6;
    rg.active = F64_GATES;   // every warp takes part in every round (a warp without a product repeats the last one and does not store)
    rg.period = 6;
    uint64_t* inbar = reinterpret_cast<uint64_t*>(gbase + 2 * 1024 * 4 + 512 * 16 + 6 * 32 * 16);   // this warp's "next input has landed"
    if (lane == 0) mbar_init(inbar, 1);
    if (threadIdx.x == 0) {
        for (int s = 0; s < F64_RING; s++) { mbar_init(rg.full + s, 1); mbar_init(rg.empty + s, F64_GATES); rg.left[s] = 0; }
    }
    if (lane == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    {
        double* t = reinterpret_cast<double*>(tab);
        for (int k = threadIdx.x; k < F64_FWDB_ROWS * 64; k += blockDim.x) t[k] = g_f64_fwdB[k];
        for (int k = threadIdx.x; k < F64_INVA_ROWS * 64; k += blockDim.x) t[F64_FWDB_ROWS * 64 + k] = g_f64_invA[k];
        for (int k = threadIdx.x; k < F64_UNTW_ROWS * 64; k += blockDim.x) t[(F64_FWDB_ROWS + F64_INVA_ROWS) * 64 + k] = g_f64_untw[k];
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (long n = 0; n < F64_RING && n < rg.total; n++)
            bulk_fetch(rg.slot + (size_t)n * F64_SLOT_ELEMS, rg.key + (size_t)(n % 6) * F64_SLOT_ELEMS, F64_CHUNK_BYTES, rg.full + n);
    long n = 0;
    // plain products (no rep_0): the result goes straight to global memory from the inverse transforms, so the input buffer is free
    // once both polynomials' digit planes exist -- the NEXT product's input is fetched into it by one bulk (TMA) copy under the rest
    // of this product (three forward and two inverse transforms)
    const bool plain = a.trlwe_in0 == nullptr;
    bool have_input = false;
    uint32_t in_par = 0;
#pragma unroll 1
    for (long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const long want = grp * F64_GATES + gl;
        const long g = want < a.B ? want : a.B - 1;
        if (have_input) {
            mbar_wait(inbar, in_par);
            in_par ^= 1u;
        } else {   // 16 independent 16-byte loads per lane (one memory latency per product, not one per 128-byte row)
            const uint4* src = reinterpret_cast<const uint4*>(a.trlwe_in + (size_t)g * 2048) + lane;
            uint4* dst = reinterpret_cast<uint4*>(acc) + lane;
            uint4 v[16];
#pragma unroll
            for (int q = 0; q < 16; q++) v[q] = src[32 * q];
            if (a.trlwe_in0) {   // cmux: rep_1 - rep_0
                const uint4* sub = reinterpret_cast<const uint4*>(a.trlwe_in0 + (size_t)g * 2048) + lane;
#pragma unroll
                for (int q = 0; q < 16; q++) { const uint4 w = sub[32 * q]; v[q].x -= w.x; v[q].y -= w.y; v[q].z -= w.z; v[q].w -= w.w; }
            }
#pragma unroll
            for (int q = 0; q < 16; q++) dst[32 * q] = v[q];
        }
        __syncwarp();
        cd s0[16], s1[16];
#pragma unroll
        for (int k = 0; k < 16; k++) { s0[k].re = 0.0; s0[k].im = 0.0; s1[k].re = 0.0; s1[k].im = 0.0; }
#pragma unroll 1
        for (int pw = 0; pw < 2; pw++) {
            {
                uint32_t u[32];
                t2_u<false>(lane, acc + pw * 1024, 0u, a.mask, u);
                u4 re, im;
                f64_pack_plane<0>(u, re, im);
                D[0 * 32 + lane] = make_uint4(re.x, re.y, re.z, re.w); D[1 * 32 + lane] = make_uint4(im.x, im.y, im.z, im.w);
                f64_pack_plane<1>(u, re, im);
                D[2 * 32 + lane] = make_uint4(re.x, re.y, re.z, re.w); D[3 * 32 + lane] = make_uint4(im.x, im.y, im.z, im.w);
                f64_pack_plane<2>(u, re, im);
                D[4 * 32 + lane] = make_uint4(re.x, re.y, re.z, re.w); D[5 * 32 + lane] = make_uint4(im.x, im.y, im.z, im.w);
            }
            if (pw == 1) {
                const long nxt = want + (long)gridDim.x * F64_GATES;
                have_input = plain && nxt < a.B;
                __syncwarp();   // every lane has taken its source words
                if (have_input && lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    bulk_fetch(acc, a.trlwe_in + (size_t)nxt * 2048, 2048u * 4u, inbar);
                }
            }
#pragma unroll 1
            for (int dw = 0; dw < 3; dw++) {
                cd x[16], y[16];
                {
                    const uint4 a4 = D[(2 * dw) * 32 + lane], b4 = D[(2 * dw + 1) * 32 + lane];
                    u4 re, im;
                    re.x = a4.x; re.y = a4.y; re.z = a4.z; re.w = a4.w; im.x = b4.x; im.y = b4.y; im.z = b4.z; im.w = b4.w;
                    f64_digits(re, im, x);
                }
                f64_forward(lane, x, S, tb, y);
                f64_with_chunk(rg, n, lane, [&](const cd16* k) {
                    f64_mac(lane, y, k, s0);
                    f64_mac(lane, y, k + F64_CHUNK_ELEMS, s1);
                });
                n++;
            }
        }
        uint32_t* ao = (plain && want < a.B) ? a.trlwe_out + (size_t)g * 2048 : acc;   // (a warp without a product has no next input either)
        f64_inverse_acc<true>(lane, s0, S, ta, ut, ao);
        f64_inverse_acc<true>(lane, s1, S, ta, ut, ao + 1024);
        __syncwarp();
        if (!plain && want < a.B) {
            uint4* dst = reinterpret_cast<uint4*>(a.trlwe_out + (size_t)g * 2048) + lane;
            const uint4* res = reinterpret_cast<const uint4*>(acc) + lane;
            uint4 v[16];
#pragma unroll
            for (int q = 0; q < 16; q++) v[q] = res[32 * q];
            if (a.trlwe_in0) {   // cmux: ... + rep_0
                const uint4* add = reinterpret_cast<const uint4*>(a.trlwe_in0 + (size_t)g * 2048) + lane;
#pragma unroll
                for (int q = 0; q < 16; q++) { const uint4 w = add[32 * q]; v[q].x += w.x; v[q].y += w.y; v[q].z += w.z; v[q].w += w.w; }
            }
#pragma unroll
            for (int q = 0; q < 16; q++) dst[32 * q] = v[q];
        }
        __syncwarp();
    }
}

// =====================================================================================================
// K5FXI: external products / cmux with ONE TRGSW PER ITEM in the FFT64 arithmetic (TRGSWRepF::cross / cmux on a fresh TRGSW,
// trgsw.rs:264-322; BASELINE config 3):  out[g] = TRGSW[g] (x) (in[g] - in0[g]) + in0[g].
// There is no key to share: a warp transforms the twelve polynomials of its item's TRGSW itself, straight from their torus words
// in global memory (18 forward + 2 inverse transforms per product instead of 8).  The spectrum of a digit polynomial waits in
// shared memory (on the accumulator's space: both polynomials' digit planes are taken first) while the two key polynomials of
// its row are transformed in the registers and multiplied in; 1 / 512 and the byte scale of the digits go onto the sums.
// =====================================================================================================
constexpr int XPI_WARPS = 8;
constexpr int XPI_WARP_BYTES = 2 * 1024 * 4 /*input / digit spectrum / result*/ + 512 * 16 /*transpose scratch*/ + 12 * 32 * 16 /*digit planes of both polynomials*/;
constexpr int XPI_SMEM_BYTES = (F64_TAB_ELEMS + F64_UNTW_ROWS * 32) * 16 + XPI_WARPS * XPI_WARP_BYTES;
static_assert(XPI_SMEM_BYTES <= 227 * 1024, "per-item external product: shared memory of one SM");
__global__ void __launch_bounds__(XPI_WARPS * 32, 1) external_product_item_f64_kernel(const BrArgs a, const uint32_t* __restrict__ trgsw /*[B][6][2][1024]*/) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd16* tab = reinterpret_cast<cd16*>(smem_raw);
    const cd16* tb = tab;
    const cd16* ta = tab + F64_FWDB_ROWS * 32;
    const cd16* ut = tab + F64_TAB_ELEMS;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* wbase = smem_raw + (F64_TAB_ELEMS + F64_UNTW_ROWS * 32) * 16 + (size_t)w * XPI_WARP_BYTES;
    uint32_t* acc = reinterpret_cast<uint32_t*>(wbase);
    cd16* Y = reinterpret_cast<cd16*>(wbase) + lane;   // the digit spectrum, lane-private, on the accumulator's space
    cd16* S = reinterpret_cast<cd16*>(wbase + 2 * 1024 * 4);
    uint4* D = reinterpret_cast<uint4*>(wbase + 2 * 1024 * 4 + 512 * 16);
    {
        double* t = reinterpret_cast<double*>(tab);
        for (int k = threadIdx.x; k < F64_FWDB_ROWS * 64; k += blockDim.x) t[k] = g_f64_fwdB[k];
        for (int k = threadIdx.x; k < F64_INVA_ROWS * 64; k += blockDim.x) t[F64_FWDB_ROWS * 64 + k] = g_f64_invA[k];
        for (int k = threadIdx.x; k < F64_UNTW_ROWS * 64; k += blockDim.x) t[(F64_FWDB_ROWS + F64_INVA_ROWS) * 64 + k] = g_f64_untw[k];
    }
    __syncthreads();
#pragma unroll 1
    for (long g = (long)blockIdx.x * XPI_WARPS + w; g < a.B; g += (long)gridDim.x * XPI_WARPS) {
        {
            const uint4* src = reinterpret_cast<const uint4*>(a.trlwe_in + (size_t)g * 2048) + lane;
            uint4* dst = reinterpret_cast<uint4*>(acc) + lane;
            uint4 v[16];
#pragma unroll
            for (int q = 0; q < 16; q++) v[q] = src[32 * q];
            if (a.trlwe_in0) {   // cmux: rep_1 - rep_0
                const uint4* sub = reinterpret_cast<const uint4*>(a.trlwe_in0 + (size_t)g * 2048) + lane;
#pragma unroll
                for (int q = 0; q < 16; q++) { const uint4 t = sub[32 * q]; v[q].x -= t.x; v[q].y -= t.y; v[q].z -= t.z; v[q].w -= t.w; }
            }
#pragma unroll
            for (int q = 0; q < 16; q++) dst[32 * q] = v[q];
        }
        __syncwarp();
#pragma unroll 1
        for (int pw = 0; pw < 2; pw++) {   // the digit planes of both polynomials: the accumulator's space is free afterwards
            uint32_t u[32];
            t2_u<false>(lane, acc + pw * 1024, 0u, a.mask, u);
            u4 re, im;
            uint4* Dp = D + pw * 6 * 32 + lane;
            f64_pack_plane<0>(u, re, im);
            Dp[0 * 32] = make_uint4(re.x, re.y, re.z, re.w); Dp[1 * 32] = make_uint4(im.x, im.y, im.z, im.w);
            f64_pack_plane<1>(u, re, im);
            Dp[2 * 32] = make_uint4(re.x, re.y, re.z, re.w); Dp[3 * 32] = make_uint4(im.x, im.y, im.z, im.w);
            f64_pack_plane<2>(u, re, im);
            Dp[4 * 32] = make_uint4(re.x, re.y, re.z, re.w); Dp[5 * 32] = make_uint4(im.x, im.y, im.z, im.w);
        }
        __syncwarp();   // every lane has read its source words (other lanes' rows of acc): Y may overwrite them
        cd s0[16], s1[16];
#pragma unroll
        for (int k = 0; k < 16; k++) { s0[k].re = 0.0; s0[k].im = 0.0; s1[k].re = 0.0; s1[k].im = 0.0; }
        const uint32_t* item = trgsw + (size_t)g * (12 * 1024);
        // every key polynomial (4 KB = 32 lines of 128 bytes, one per lane) is requested into L1 while the previous transform runs
        auto prefetch_poly = [&](int q) { asm volatile("prefetch.global.L1 [%0];" ::"l"(item + (size_t)q * 1024 + lane * 32)); };
        prefetch_poly(0);
#pragma unroll 1
        for (int j = 0; j < 6; j++) {
            cd x[16], y[16];
            {
                const uint4 a4 = D[(2 * j) * 32 + lane], b4 = D[(2 * j + 1) * 32 + lane];
                u4 re, im;
                re.x = a4.x; re.y = a4.y; re.z = a4.z; re.w = a4.w; im.x = b4.x; im.y = b4.y; im.z = b4.z; im.w = b4.w;
                f64_digits(re, im, x);
            }
            f64_forward(lane, x, S, tb, y);
#pragma unroll
            for (int k = 0; k < 16; k++) { cd16 v; v.re = y[k].re; v.im = y[k].im; Y[32 * k] = v; }
            f64_key_input(lane, item + (size_t)(2 * j) * 1024, x);
            prefetch_poly(2 * j + 1);
            f64_forward(lane, x, S, tb, y);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const cd16 d = Y[32 * k];
                s0[k].re = F_FMA(d.re, y[k].re, F_FMA(-d.im, y[k].im, s0[k].re));
                s0[k].im = F_FMA(d.re, y[k].im, F_FMA(d.im, y[k].re, s0[k].im));
            }
            f64_key_input(lane, item + (size_t)(2 * j + 1) * 1024, x);
            if (j < 5) prefetch_poly(2 * j + 2);
            f64_forward(lane, x, S, tb, y);
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const cd16 d = Y[32 * k];
                s1[k].re = F_FMA(d.re, y[k].re, F_FMA(-d.im, y[k].im, s1[k].re));
                s1[k].im = F_FMA(d.re, y[k].im, F_FMA(d.im, y[k].re, s1[k].im));
            }
        }
#pragma unroll
        for (int k = 0; k < 16; k++) {   // 1 / 512 of the inverse transform and 1 / 4 of the digit bytes: an exact power of two
            s0[k].re = F_MUL(s0[k].re, F64_KEY_SCALE); s0[k].im = F_MUL(s0[k].im, F64_KEY_SCALE);
            s1[k].re = F_MUL(s1[k].re, F64_KEY_SCALE); s1[k].im = F_MUL(s1[k].im, F64_KEY_SCALE);
        }
        __syncwarp();
        f64_inverse_acc<true>(lane, s0, S, ta, ut, acc);
        f64_inverse_acc<true>(lane, s1, S, ta, ut, acc + 1024);
        __syncwarp();
        {
            uint4* dst = reinterpret_cast<uint4*>(a.trlwe_out + (size_t)g * 2048) + lane;
            const uint4* res = reinterpret_cast<const uint4*>(acc) + lane;
            uint4 v[16];
#pragma unroll
            for (int q = 0; q < 16; q++) v[q] = res[32 * q];
            if (a.trlwe_in0) {   // cmux: ... + rep_0
                const uint4* add = reinterpret_cast<const uint4*>(a.trlwe_in0 + (size_t)g * 2048) + lane;
#pragma unroll
                for (int q = 0; q < 16; q++) { const uint4 t = add[32 * q]; v[q].x += t.x; v[q].y += t.y; v[q].z += t.z; v[q].w += t.w; }
            }
#pragma unroll
            for (int q = 0; q < 16; q++) dst[32 * q] = v[q];
        }
        __syncwarp();
    }
}

// =====================================================================================================
// K7F: exact negacyclic product a * d mod (X^N + 1, 2^32) in the FFT64 arithmetic (Polynomial::fft_cross, math.rs:337-347;
// Spqlios_poly_mul, spqlios-wrapper.cpp:38-53; BASELINE config 3).  a = torus words taken as centred 32-bit integers, |d| <= 192:
// a coefficient of the product is below 1024 * 2^31 * 192 < 2^49, as in the external product.  One product per warp (two forward
// transforms, a pointwise product, one inverse), persistent CTAs of twelve warps; inputs are read and the result is written
// straight from / to global memory (coalesced 128-byte rows), shared memory holds only the twiddle tables and the transpose scratch.
// =====================================================================================================
constexpr int PMF_WARPS = 12;
constexpr int PMF_SMEM_BYTES = (F64_TAB_ELEMS + F64_UNTW_ROWS * 32) * 16 + 2 * PMF_WARPS * 512 * 16;   // per warp: transpose scratch + the first spectrum
__global__ void __launch_bounds__(PMF_WARPS * 32, 1) polymul_f64_kernel(const uint32_t* __restrict__ a, const int32_t* __restrict__ d,
                                                                       uint32_t* __restrict__ out, long B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd16* tab = reinterpret_cast<cd16*>(smem_raw);
    const cd16* tb = tab;
    const cd16* ta = tab + F64_FWDB_ROWS * 32;
    const cd16* ut = tab + F64_TAB_ELEMS;
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    cd16* S = tab + F64_TAB_ELEMS + F64_UNTW_ROWS * 32 + (size_t)w * 1024;
    cd16* Y = S + 512 + lane;   // the first spectrum waits here while the second transform has the registers (168 per thread at twelve warps)
    {
        double* t = reinterpret_cast<double*>(tab);
        for (int k = threadIdx.x; k < F64_FWDB_ROWS * 64; k += blockDim.x) t[k] = g_f64_fwdB[k];
        for (int k = threadIdx.x; k < F64_INVA_ROWS * 64; k += blockDim.x) t[F64_FWDB_ROWS * 64 + k] = g_f64_invA[k];
        for (int k = threadIdx.x; k < F64_UNTW_ROWS * 64; k += blockDim.x) t[(F64_FWDB_ROWS + F64_INVA_ROWS) * 64 + k] = g_f64_untw[k];
    }
    __syncthreads();
#pragma unroll 1
    for (long g = (long)blockIdx.x * PMF_WARPS + w; g < B; g += (long)gridDim.x * PMF_WARPS) {
        cd x[16], y[16];
        f64_key_input(lane, a + (size_t)g * 1024, x);
        const int32_t* dp = d + (size_t)g * 1024;
        {   // the second operand into L1 and the next product's operands into L2 while the first transform runs (one 128-byte line per lane)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(dp + lane * 32));
            const long gn = g + (long)gridDim.x * PMF_WARPS;
            if (gn < B) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(a + (size_t)gn * 1024 + lane * 32));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(d + (size_t)gn * 1024 + lane * 32));
            }
        }
        f64_forward(lane, x, S, tb, y);
#pragma unroll
        for (int k = 0; k < 16; k++) { cd16 v; v.re = y[k].re; v.im = y[k].im; Y[32 * k] = v; }
#pragma unroll
        for (int r = 0; r < 16; r++) { x[r].re = (double)dp[32 * r + lane]; x[r].im = (double)dp[512 + 32 * r + lane]; }
        f64_forward(lane, x, S, tb, y);
#pragma unroll
        for (int k = 0; k < 16; k++) {   // pointwise product, with the 1/512 of the inverse transform (exact power of two)
            const cd16 ya = Y[32 * k];
            const double pr = F_FMA(ya.re, y[k].re, -F_MUL(ya.im, y[k].im));
            const double pi = F_FMA(ya.re, y[k].im, F_MUL(ya.im, y[k].re));
            y[k].re = F_MUL(pr, 1.0 / 512); y[k].im = F_MUL(pi, 1.0 / 512);
        }
        f64_inverse_acc<true>(lane, y, S, ta, ut, out + (size_t)g * 1024);
    }
}
