// blind_rotate_f64t.cuh -- K5FT, the FFT64 throughput blind rotation with its per-gate state in TENSOR MEMORY (included by engine.cu).
//   gate pre-combination + 635 x CMUX + sample extract (tfhe.rs:27-113, trgsw.rs:264-322, trlwe.rs:110-121)
// K5F (blind_rotate_f64.cuh) keeps the two output spectra of a gate in 128 registers, which caps an SM at eight warps -- two per
// scheduler -- and leaves the FP64 pipe idle half of the time.  Blackwell's tensor memory (256 KB per SM,
// its own load / store path: tcgen05.ld / tcgen05.st) is thread-private storage when it is addressed with the .32x32b shape
// (thread t of warp w owns lane 32 (w % 4) + t, any columns).  Here it holds, per gate: the two output spectra (2 x 64 words per
// thread), the three digit byte planes (24 words) and the rounded LWE mask abar (10 words).  A warp then needs about 130
// registers and 16 KB of shared memory (accumulator polynomials + transpose scratch): TWELVE gates per SM, three warps per
// scheduler.  No tensor-core instruction is involved; tensor memory is used as a second register file.
// Everything else (transform, key ring, exact rounding) is K5F's; results are bit-identical.
// MEASURED AND NOT ADOPTED (opt-in: TFHE_B200_F64_TMEM=1): 1776 gates in 18.2 ms = 97.6 k gates/s against 108 k for K5F.  The third
// warp per scheduler does not raise the issue rate (52 % against 48 %): a DFMA occupies the issue port for two cycles
// (profiles/r02_dfma_mix.json), so the kernel is bound by issue slots, and this variant executes more instructions per gate
// (tensor-memory loads, stores and waits; a 3 x 8 KB ring; the untwist table through L1).  Kept as the evidence, and as the
// template for using tensor memory as thread-private storage.
#pragma once
#include "blind_rotate_f64.cuh"

constexpr int F64T_GATES = 12;
constexpr int F64T_RING = 3;                               // ring slots of one (row, output) key polynomial: 8 KB each
constexpr int F64T_SLOT_BYTES = (int)(F64_CHUNK_ELEMS * sizeof(cd16));
constexpr int F64T_GATE_SMEM_BYTES = 2 * 1024 * 4 /*acc*/ + 512 * 16 /*transpose scratch*/;
constexpr int F64T_SHARED_BYTES = (F64_TAB_ELEMS * 16 + F64T_RING * F64T_SLOT_BYTES + 2 * F64T_RING * 8 + F64T_RING * 4 + 4 + 15) / 16 * 16;
constexpr size_t f64t_smem_bytes() { return (size_t)F64T_SHARED_BYTES + (size_t)F64T_GATES * F64T_GATE_SMEM_BYTES; }
static_assert(f64t_smem_bytes() <= 227 * 1024, "twelve gates and the key ring must fit the shared memory of one SM");
// tensor-memory columns of one warp slot (three warps share a 32-lane quadrant)
constexpr uint32_t F64T_COL_S0 = 0, F64T_COL_S1 = 64, F64T_COL_PLANES = 128, F64T_COL_ABAR = 152, F64T_COLS_PER_WARP = 168;
static_assert(3 * F64T_COLS_PER_WARP <= 512, "tensor memory has 512 columns");

__device__ __forceinline__ void tm_ld16(uint32_t addr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
                   "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(addr));
}
__device__ __forceinline__ void tm_st16(uint32_t addr, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(addr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]),
                 "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tm_ld8(uint32_t addr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(addr));
}
__device__ __forceinline__ void tm_st8(uint32_t addr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]),
                 "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ uint32_t tm_ld1(uint32_t addr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(addr));
    return r;
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// one output spectrum (16 complex values per lane = 64 words) between tensor memory and registers
__device__ __forceinline__ void f64t_load_spectrum(uint32_t taddr, cd (&s)[16]) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t r[16];
        tm_ld16(taddr + 16 * c, r);
        tm_wait_ld();
#pragma unroll
        for (int m = 0; m < 4; m++) {
            s[4 * c + m].re = __hiloint2double((int)r[4 * m + 1], (int)r[4 * m]);
            s[4 * c + m].im = __hiloint2double((int)r[4 * m + 3], (int)r[4 * m + 2]);
        }
    }
}
// acc (tensor memory) += y * key, four complex values at a time; FIRST: acc = y * key
template <bool FIRST>
__device__ __forceinline__ void f64t_mac(int lane, const cd (&y)[16], const cd16* key, uint32_t taddr) {
    uint32_t r[4][16];
    if (!FIRST) {
#pragma unroll
        for (int c = 0; c < 4; c++) tm_ld16(taddr + 16 * c, r[c]);
        tm_wait_ld();
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const int k = 4 * c + m;
            const cd16 w = key[k * 32 + lane];
            double ar, ai;
            if (FIRST) {
                ar = F_FMA(y[k].re, w.re, -F_MUL(y[k].im, w.im));
                ai = F_FMA(y[k].re, w.im, F_MUL(y[k].im, w.re));
            } else {
                ar = __hiloint2double((int)r[c][4 * m + 1], (int)r[c][4 * m]);
                ai = __hiloint2double((int)r[c][4 * m + 3], (int)r[c][4 * m + 2]);
                ar = F_FMA(y[k].re, w.re, F_FMA(-y[k].im, w.im, ar));
                ai = F_FMA(y[k].re, w.im, F_FMA(y[k].im, w.re, ai));
            }
            r[c][4 * m] = (uint32_t)__double2loint(ar); r[c][4 * m + 1] = (uint32_t)__double2hiint(ar);
            r[c][4 * m + 2] = (uint32_t)__double2loint(ai); r[c][4 * m + 3] = (uint32_t)__double2hiint(ai);
        }
        tm_st16(taddr + 16 * c, r[c]);
    }
}
// the untwist table read through L1 (there is no shared memory left for it)
__device__ __forceinline__ void f64t_untwist_round(int lane, const cd (&w)[16], uint32_t (&lo)[16], uint32_t (&hi)[16]) {
    const double2* ut = reinterpret_cast<const double2*>(g_f64_untw) + lane;
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const double2 t = __ldg(ut + r * 32);
        const double zr = F_FMA(w[r].re, t.x, -F_MUL(w[r].im, t.y));
        const double zi = F_FMA(w[r].re, t.y, F_MUL(w[r].im, t.x));
        lo[r] = f64_low_word(F_ADD(zr, F64_ROUND_MAGIC));
        hi[r] = f64_low_word(F_ADD(zi, F64_ROUND_MAGIC));
    }
}
__device__ __forceinline__ void f64t_inverse_acc(int lane, cd (&sp)[16], cd16* S, const cd16* ta, uint32_t* ao) {
    cd v[16];
    {
        f64_inv_low(sp);
        cd send[8], recv[8];
        f64_x_send(lane, sp, send);
        f64_exchange(send, recv);
        f64_inv_x_bfly(lane, sp, recv, v);
    }
    F64TwA tw;
    f64_inv_twA(lane, ta, tw);
    f64_t2_store(lane, v, S);
    __syncwarp();
    f64_t2_load(lane, S, v);
    __syncwarp();
    f64_inv_passA(v, tw);
    uint32_t lo[16], hi[16];
    f64t_untwist_round(lane, v, lo, hi);
#pragma unroll
    for (int r = 0; r < 16; r++) { ao[32 * r + lane] += lo[r]; ao[512 + 32 * r + lane] += hi[r]; }
}

struct F64TRing {
    cd16* slot;
    uint64_t* full;
    uint64_t* empty;
    uint32_t* left;
    const cd16* key;
    long total;
    int active;
};
__device__ __forceinline__ const cd16* f64t_chunk_wait(const F64TRing& rg, long n) {
    const int s = (int)(n % F64T_RING);
    mbar_wait(rg.full + s, (uint32_t)((n / F64T_RING) & 1));
    return rg.slot + (size_t)s * F64_CHUNK_ELEMS;
}
// hand the slot of chunk n back; the LAST warp to leave it requests chunk n + F64T_RING
__device__ __forceinline__ void f64t_chunk_release(const F64TRing& rg, long n, int lane) {
    __syncwarp();
    if (lane == 0) {
        const int s = (int)(n % F64T_RING);
        mbar_arrive(rg.empty + s);
        if (atomicAdd(rg.left + s, 1u) == (uint32_t)(rg.active - 1)) {
            rg.left[s] = 0;
            if (n + F64T_RING < rg.total) {
                mbar_wait(rg.empty + s, (uint32_t)((n / F64T_RING) & 1));
                bulk_fetch(rg.slot + (size_t)s * F64_CHUNK_ELEMS, rg.key + (size_t)(n + F64T_RING) * F64_CHUNK_ELEMS, F64T_SLOT_BYTES, rg.full + s);
            }
        }
    }
}

__global__ void __launch_bounds__(F64T_GATES * 32, 1) blind_rotate_f64t_kernel(const BrArgs a, const cd16* __restrict__ key) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd16* tab = reinterpret_cast<cd16*>(smem_raw);
    const cd16* tb = tab;                               // forward pass B / exchange twiddles
    const cd16* ta = tab + F64_FWDB_ROWS * 32;          // inverse stages 5..8
    F64TRing rg;
    rg.slot = tab + F64_TAB_ELEMS;
    rg.full = reinterpret_cast<uint64_t*>(rg.slot + (size_t)F64T_RING * F64_CHUNK_ELEMS);
    rg.empty = rg.full + F64T_RING;
    rg.left = reinterpret_cast<uint32_t*>(rg.empty + F64T_RING);
    uint32_t* tmem_base_slot = rg.left + F64T_RING;
    const int gl = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* gbase = smem_raw + F64T_SHARED_BYTES + (size_t)gl * F64T_GATE_SMEM_BYTES;
    uint32_t* acc = reinterpret_cast<uint32_t*>(gbase);
    cd16* S = reinterpret_cast<cd16*>(gbase + 2 * 1024 * 4);

    const long cta = blockIdx.x;
    const long first = cta * a.cta_base + (cta < a.cta_rem ? cta : a.cta_rem);
    const int cnt = a.cta_base + (cta < a.cta_rem ? 1 : 0);
    const bool active = gl < cnt;
    const long gate = active ? first + gl : a.B - 1;
    const int nsteps = a.nsteps;
    rg.key = key;
    rg.total = (long)nsteps * 12;
    rg.active = cnt;

    if (gl == 0) {   // tensor memory: all 512 columns of this SM (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"l"((uint64_t)__cvta_generic_to_shared(tmem_base_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < F64T_RING; s++) { mbar_init(rg.full + s, 1); mbar_init(rg.empty + s, cnt); rg.left[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        double* t = reinterpret_cast<double*>(tab);
        for (int k = threadIdx.x; k < F64_FWDB_ROWS * 64; k += blockDim.x) t[k] = g_f64_fwdB[k];
        for (int k = threadIdx.x; k < F64_INVA_ROWS * 64; k += blockDim.x) t[F64_FWDB_ROWS * 64 + k] = g_f64_invA[k];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // tables, mbarriers, tensor-memory base
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *tmem_base_slot + ((uint32_t)(32 * (gl & 3)) << 16) + F64T_COLS_PER_WARP * (uint32_t)(gl >> 2);
    if (threadIdx.x == 0)
        for (long n = 0; n < F64T_RING && n < rg.total; n++)
            bulk_fetch(rg.slot + (size_t)n * F64_CHUNK_ELEMS, rg.key + (size_t)n * F64_CHUNK_ELEMS, F64T_SLOT_BYTES, rg.full + n);

    if (active) {
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
            {   // abar_i, rounded (tfhe.rs:107-108), two per word: lane l keeps the steps i = 64 q + 2 l, 64 q + 2 l + 1 in column q
                uint32_t ab[16];
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    const int i0 = 64 * q + 2 * lane;
                    const uint32_t lo = (q < 10 && i0 < LWE_N) ? ((lin[1 + i0] + (1u << 20)) >> 21) & 0xFFFFu : 0u;
                    const uint32_t hi = (q < 10 && i0 + 1 < LWE_N) ? ((lin[2 + i0] + (1u << 20)) >> 21) & 0xFFFFu : 0u;
                    ab[q] = lo | (hi << 16);
                }
                tm_st16(tm + F64T_COL_ABAR, ab);
            }
            const uint32_t bbar = lin[0] >> 21;                                                             // floor
            const uint32_t nrot = (2048u - bbar) & 2047u;   // acc_0 = X^{-bbar} * (mu, ..., mu ; 0)
            __syncwarp();
            for (int k = lane; k < 1024; k += 32) {
                const bool neg = ((uint32_t)k < (nrot & 1023u)) != (nrot >= 1024u);
                acc[k] = neg ? 0u - a.mu : a.mu;
                acc[1024 + k] = 0;
            }
            tm_wait_st();
            __syncwarp();
        }

        // ---- 635 x CMUX ----
        long n = 0;
#pragma unroll 1
        for (int i = 0; i < nsteps; i++) {
            uint32_t ab;
            {
                const uint32_t wv = tm_ld1(tm + F64T_COL_ABAR + (uint32_t)(i >> 6));
                tm_wait_ld();
                const uint32_t pr = __shfl_sync(0xffffffffu, wv, (i >> 1) & 31);
                ab = (i & 1) ? pr >> 16 : pr & 0xFFFFu;
            }
#pragma unroll 1
            for (int pw = 0; pw < 2; pw++) {
                {   // masked source words of polynomial pw: lane-private, parked in tensor memory as three byte planes
                    uint32_t u[32];
                    t2_u<true>(lane, acc + pw * 1024, ab, a.mask, u);
                    u4 re, im;
                    uint32_t p[8];
                    f64_pack_plane<0>(u, re, im);
                    p[0] = re.x; p[1] = re.y; p[2] = re.z; p[3] = re.w; p[4] = im.x; p[5] = im.y; p[6] = im.z; p[7] = im.w;
                    tm_st8(tm + F64T_COL_PLANES, p);
                    f64_pack_plane<1>(u, re, im);
                    p[0] = re.x; p[1] = re.y; p[2] = re.z; p[3] = re.w; p[4] = im.x; p[5] = im.y; p[6] = im.z; p[7] = im.w;
                    tm_st8(tm + F64T_COL_PLANES + 8, p);
                    f64_pack_plane<2>(u, re, im);
                    p[0] = re.x; p[1] = re.y; p[2] = re.z; p[3] = re.w; p[4] = im.x; p[5] = im.y; p[6] = im.z; p[7] = im.w;
                    tm_st8(tm + F64T_COL_PLANES + 16, p);
                    tm_wait_st();
                }
#pragma unroll 1
                for (int dw = 0; dw < 3; dw++) {
                    cd x[16], y[16];
                    {
                        uint32_t p[8];
                        tm_ld8(tm + F64T_COL_PLANES + 8 * (uint32_t)dw, p);
                        tm_wait_ld();
                        u4 re, im;
                        re.x = p[0]; re.y = p[1]; re.z = p[2]; re.w = p[3]; im.x = p[4]; im.y = p[5]; im.z = p[6]; im.w = p[7];
                        f64_digits(re, im, x);
                    }
                    f64_forward(lane, x, S, tb, y);
                    const bool first_digit = (pw | dw) == 0;
                    {
                        const cd16* k = f64t_chunk_wait(rg, n);
                        if (first_digit) f64t_mac<true>(lane, y, k, tm + F64T_COL_S0); else f64t_mac<false>(lane, y, k, tm + F64T_COL_S0);
                        f64t_chunk_release(rg, n, lane);
                    }
                    {
                        const cd16* k = f64t_chunk_wait(rg, n + 1);
                        if (first_digit) f64t_mac<true>(lane, y, k, tm + F64T_COL_S1); else f64t_mac<false>(lane, y, k, tm + F64T_COL_S1);
                        f64t_chunk_release(rg, n + 1, lane);
                    }
                    n += 2;
                    tm_wait_st();
                }
            }
#pragma unroll 1
            for (int o = 0; o < 2; o++) {
                cd sp[16];
                f64t_load_spectrum(tm + (o ? F64T_COL_S1 : F64T_COL_S0), sp);
                f64t_inverse_acc(lane, sp, S, ta, acc + o * 1024);
            }
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
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // every warp is done with its tensor-memory columns
    if (gl == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(*tmem_base_slot) : "memory");
}
