// blind_rotate_f64w2.cuh -- K5F2, FFT64 throughput blind rotation with one gate on TWO warps (included by engine.cu only).
//   gate pre-combination + 635 x CMUX + sample extract (tfhe.rs:27-113, trgsw.rs:264-322, trlwe.rs:110-121)
// K5F (one warp per gate, 16 values per lane) is bound by issue slots, and a 16-values-per-lane transform has one stage that
// pairs values of two lanes (shuffles and selects; when this kernel was written K5F paid 128 of them per transform, it now takes
// the forward's on the load side of its transpose).  With 8 values per thread on two warps a transform is three
// radix-8 passes and two transposes through shared memory (the transform of the latency kernel, blind_rotate_f64l2.cuh): no such
// stage, the per-thread twiddles are 4 + 4 loads, and the two output spectra of a gate are 64 registers per thread instead of
// 128 -- so G gates = 2 G warps fit an SM with more than two warps per scheduler.
// Per step and gate: both polynomials' masked source words (16 per thread and polynomial, thread-private: thread t needs the
// coefficients t + 64 e and 512 + t + 64 e) are packed into three digit byte planes that stay in REGISTERS (12); six forward
// transforms, each multiplied into both output spectra from the key ring (K5F's ring; the key in the [register 8][thread 64]
// layout); two inverse transforms; exact rounding; accumulate.  Barriers: two 64-thread named barriers per transform, one per step.
// Measured (profiles/README.md): 94.5 k gates/s in the kernel at six gates per SM against 108.4 k for K5F -- 7 891 instructions per
// gate and CMUX instead of 8 251, but 64 % of the issue slots used instead of 75 % (168 registers with 37 spilled words, barrier and
// shared-memory latency on three warps per scheduler); releasing the key slot once per gate after the gate's barrier and starting
// the sums with a multiplication (-190 instructions) made it 1 % slower.  Opt-in: TFHE_B200_F64_KERNEL=w2.
#pragma once
#include "blind_rotate_f64l2.cuh"

#if !defined(F64W2_GATES_DEF)
#define F64W2_GATES_DEF 6
#endif
constexpr int F64W2_GATES = F64W2_GATES_DEF;
constexpr int F64W2_THREADS = F64W2_GATES * 64;
constexpr int F64W2_GATE_SMEM_BYTES = 2 * 1024 * 4 /*acc*/ + 2 * 512 * 16 /*transpose buffers A, B*/ + 640 * 2 /*abar*/;
constexpr int F64W2_SHARED_BYTES = (F64L2_TAB_ELEMS * 16 + F64_RING * F64_CHUNK_BYTES + 2 * F64_RING * 8 + F64_RING * 4 + 15) / 16 * 16;
constexpr size_t f64w2_smem_bytes() { return (size_t)F64W2_SHARED_BYTES + (size_t)F64W2_GATES * F64W2_GATE_SMEM_BYTES; }
static_assert(f64w2_smem_bytes() <= 227 * 1024, "the gates and the key ring must fit the shared memory of one SM");
static_assert(F64W2_GATES <= 15, "one named barrier per gate");

// 16 masked source words of one polynomial (thread t: coefficients t + 64 e -> re[e], 512 + t + 64 e -> im[e]) as three digit planes
// of 2 + 2 words: byte = 4 * digit (fft64.cuh)
template <int DW>
__device__ __forceinline__ void f64w2_pack(const uint32_t (&ur)[8], const uint32_t (&ui)[8], uint32_t (&p)[4]) {
    p[0] = f64_pack4<DW>(ur[0], ur[1], ur[2], ur[3]); p[1] = f64_pack4<DW>(ur[4], ur[5], ur[6], ur[7]);
    p[2] = f64_pack4<DW>(ui[0], ui[1], ui[2], ui[3]); p[3] = f64_pack4<DW>(ui[4], ui[5], ui[6], ui[7]);
}
__device__ __forceinline__ void f64w2_unpack(const uint32_t (&p)[4], cd (&x)[8]) {
    f64_unpack4(p[0], x[0].re, x[1].re, x[2].re, x[3].re); f64_unpack4(p[1], x[4].re, x[5].re, x[6].re, x[7].re);
    f64_unpack4(p[2], x[0].im, x[1].im, x[2].im, x[3].im); f64_unpack4(p[3], x[4].im, x[5].im, x[6].im, x[7].im);
}
// forward transform on the gate's two warps: x[e] = z_{t + 64 e}  ->  x[e] = spectrum position 8 t + e
__device__ __forceinline__ void f64w2_forward(int t, cd (&x)[8], cd16* bufA, cd16* bufB, const cd16* tf2, const cd16* tf3, int bar_id) {
    const int hi3 = t >> 3, lo3 = t & 7;
    cd16 w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) w[k] = tf2[k * 8 + hi3];   // requested before the first barrier: the latency is under pass 1
    l2_fwd_pass1(x);
#pragma unroll
    for (int e = 0; e < 8; e++) l2_store(bufA + t + 64 * e, x[e]);
    bar_sync(bar_id, 64);
#pragma unroll
    for (int m = 0; m < 8; m++) l2_load(bufA + 64 * hi3 + 8 * m + lo3, x[m]);
    l2_fwd_pass23(x, w);
#pragma unroll
    for (int k = 0; k < 4; k++) w[k] = tf3[k * 64 + t];
#pragma unroll
    for (int m = 0; m < 8; m++) l2_store(bufB + 64 * hi3 + 8 * m + (lo3 ^ m), x[m]);
    bar_sync(bar_id, 64);
#pragma unroll
    for (int e = 0; e < 8; e++) l2_load(bufB + 8 * t + (e ^ lo3), x[e]);
    l2_fwd_pass23(x, w);
}
// inverse transform of one output spectrum (destroyed), exact rounding, accumulate into the accumulator polynomial `ao`
__device__ __forceinline__ void f64w2_inverse_acc(int t, cd (&y)[8], cd16* bufA, cd16* bufB, const cd16* ti2, const cd16* ti3, const cd16* tut,
                                                  int bar_id, uint32_t* ao) {
    const int hi3 = t >> 3, lo3 = t & 7;
    cd16 v[4];
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = ti2[k * 8 + lo3];
    l2_inv_pass1(y);
#pragma unroll
    for (int e = 0; e < 8; e++) l2_store(bufB + 8 * t + (e ^ lo3), y[e]);
    bar_sync(bar_id, 64);
#pragma unroll
    for (int m = 0; m < 8; m++) l2_load(bufB + 64 * hi3 + 8 * m + (lo3 ^ m), y[m]);
    l2_inv_pass23(y, v);
#pragma unroll
    for (int k = 0; k < 4; k++) v[k] = ti3[k * 64 + t];
#pragma unroll
    for (int m = 0; m < 8; m++) l2_store(bufA + 64 * hi3 + 8 * m + lo3, y[m]);
    bar_sync(bar_id, 64);
#pragma unroll
    for (int e = 0; e < 8; e++) l2_load(bufA + t + 64 * e, y[e]);
    l2_inv_pass23(y, v);
#pragma unroll
    for (int e = 0; e < 8; e++) {
        const cd16 u = tut[e * 64 + t];
        const double zr = F_FMA(y[e].re, u.re, -F_MUL(y[e].im, u.im));
        const double zi = F_FMA(y[e].re, u.im, F_MUL(y[e].im, u.re));
        l2_red_add(ao + t + 64 * e, f64_low_word(F_ADD(zr, F64_ROUND_MAGIC)));
        l2_red_add(ao + 512 + t + 64 * e, f64_low_word(F_ADD(zi, F64_ROUND_MAGIC)));
    }
}

__global__ void __launch_bounds__(F64W2_THREADS, 1) blind_rotate_f64w2_kernel(const BrArgs a, const cd16* __restrict__ key /* [register 8][thread 64] layout */) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cd16* tab = reinterpret_cast<cd16*>(smem_raw);
    const cd16* tf2 = tab;
    const cd16* tf3 = tab + 32;
    const cd16* ti2 = tab + 288;
    const cd16* ti3 = tab + 320;
    const cd16* tut = tab + 576;
    F64Ring rg;
    rg.slot = tab + F64L2_TAB_ELEMS;
    rg.full = reinterpret_cast<uint64_t*>(rg.slot + (size_t)F64_RING * F64_SLOT_ELEMS);
    rg.empty = rg.full + F64_RING;
    rg.left = reinterpret_cast<uint32_t*>(rg.empty + F64_RING);
    f64_ring_addr(rg);
    const int gl = threadIdx.x >> 6, t = threadIdx.x & 63, lane = threadIdx.x & 31;
    unsigned char* gbase = smem_raw + F64W2_SHARED_BYTES + (size_t)gl * F64W2_GATE_SMEM_BYTES;
    uint32_t* acc = reinterpret_cast<uint32_t*>(gbase);
    cd16* bufA = reinterpret_cast<cd16*>(gbase + 2 * 1024 * 4);
    cd16* bufB = bufA + 512;
    uint16_t* abar = reinterpret_cast<uint16_t*>(bufB + 512);
    const int bar_id = 1 + gl;

    const long cta = blockIdx.x;
    const long first = cta * a.cta_base + (cta < a.cta_rem ? cta : a.cta_rem);
    const int cnt = a.cta_base + (cta < a.cta_rem ? 1 : 0);
    const bool active = gl < cnt;
    const long gate = active ? first + gl : a.B - 1;
    const int nsteps = a.nsteps;
    rg.key = key;
    rg.total = (long)nsteps * 6;
    rg.active = 2 * cnt;   // warps
    rg.period = 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < F64_RING; s++) { mbar_init(rg.full + s, 1); mbar_init(rg.empty + s, 2 * cnt); rg.left[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        double* d = reinterpret_cast<double*>(tab);
        for (int k = threadIdx.x; k < 64; k += blockDim.x) { d[k] = g_l2_fwd2[k]; d[2 * 288 + k] = g_l2_inv2[k]; }
        for (int k = threadIdx.x; k < 512; k += blockDim.x) { d[2 * 32 + k] = g_l2_fwd3[k]; d[2 * 320 + k] = g_l2_inv3[k]; }
        for (int k = threadIdx.x; k < 1024; k += blockDim.x) d[2 * 576 + k] = g_l2_untw[k];
    }
    // ---- prologue: gate pre-combination (tfhe.rs:27-71), rounding of (b, a) (tfhe.rs:97,107-108), acc_0 ----
    {
        uint32_t* lin = reinterpret_cast<uint32_t*>(bufA);
        const bool second = gate >= a.split;
        const long gsrc = second ? gate - a.split : gate;
        const uint32_t* q0 = second ? a.in0b : a.in0;
        const uint32_t* q1 = second ? a.in1b : a.in1;
        uint32_t k0 = (uint32_t)(second ? a.c0b : a.c0), k1 = (uint32_t)(second ? a.c1b : a.c1), kb = second ? a.cbb : a.cb;
        if (a.ops) gate_coeffs(a.ops[gate], a.mu, k0, k1, kb);
        const uint32_t* p0 = q0 + (size_t)(a.idx0 ? (long)a.idx0[gate] : gsrc) * (LWE_N + 1);
        const uint32_t* p1 = (q1 && k1 != 0) ? q1 + (size_t)(a.idx1 ? (long)a.idx1[gate] : gsrc) * (LWE_N + 1) : nullptr;
        for (int c = t; c <= LWE_N; c += 64) {
            uint32_t v = k0 * p0[c];
            if (p1) v += k1 * p1[c];
            if (c == 0) v += kb;
            lin[c] = v;
        }
        __syncthreads();
        for (int i = t; i < 640; i += 64) abar[i] = i < LWE_N ? (uint16_t)((lin[1 + i] + (1u << 20)) >> 21) : (uint16_t)0;   // round
        const uint32_t bbar = lin[0] >> 21;                                                                                // floor
        const uint32_t nrot = (2048u - bbar) & 2047u;
        for (int k = t; k < 1024; k += 64) {
            const bool neg = ((uint32_t)k < (nrot & 1023u)) != (nrot >= 1024u);
            acc[k] = neg ? 0u - a.mu : a.mu;
            acc[1024 + k] = 0;
        }
    }
    __syncthreads();   // tables, mbarriers, prologue
    if (threadIdx.x == 0)
        for (long n = 0; n < F64_RING && n < rg.total; n++)
            bulk_fetch(rg.slot + (size_t)n * F64_SLOT_ELEMS, rg.key + (size_t)n * F64_SLOT_ELEMS, F64_CHUNK_BYTES, rg.full + n);
    if (!active) return;   // gate slots without a gate leave here: every barrier below is private to one gate, the ring counts active warps

    // ---- 635 x CMUX ----
    long n = 0;
#pragma unroll 1
    for (int i = 0; i < nsteps; i++) {
        cd s0[8], s1[8];   // this thread's 8 points of the two output spectra
#pragma unroll
        for (int e = 0; e < 8; e++) { s0[e].re = 0.0; s0[e].im = 0.0; s1[e].re = 0.0; s1[e].im = 0.0; }
        const uint32_t ab = abar[i];
#pragma unroll 1
        for (int pw = 0; pw < 2; pw++) {
            uint32_t pl[3][4];   // the three digit planes of this thread's 16 source words
            {
                const uint32_t* A = acc + pw * 1024;
                uint32_t ur[8], ui[8];
#pragma unroll
                for (int e = 0; e < 8; e++) {
                    ur[e] = add_alu(rot_diff(A, (uint32_t)(t + 64 * e), ab), a.mask) ^ a.mask;
                    ui[e] = add_alu(rot_diff(A, (uint32_t)(512 + t + 64 * e), ab), a.mask) ^ a.mask;
                }
                f64w2_pack<0>(ur, ui, pl[0]); f64w2_pack<1>(ur, ui, pl[1]); f64w2_pack<2>(ur, ui, pl[2]);
            }
#pragma unroll
            for (int dw = 0; dw < 3; dw++) {
                cd x[8];
                f64w2_unpack(pl[dw], x);
                f64w2_forward(t, x, bufA, bufB, tf2, tf3, bar_id);
                f64_with_chunk(rg, n, lane, [&](const cd16* k) {
                    const cd16* k0 = k + t;
#pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const cd16 w0 = k0[e * 64], w1 = k0[F64_CHUNK_ELEMS + e * 64];
                        s0[e].re = F_FMA(x[e].re, w0.re, F_FMA(-x[e].im, w0.im, s0[e].re));
                        s0[e].im = F_FMA(x[e].re, w0.im, F_FMA(x[e].im, w0.re, s0[e].im));
                        s1[e].re = F_FMA(x[e].re, w1.re, F_FMA(-x[e].im, w1.im, s1[e].re));
                        s1[e].im = F_FMA(x[e].re, w1.im, F_FMA(x[e].im, w1.re, s1[e].im));
                    }
                });
                n++;
                bar_sync(bar_id, 64);   // both warps have read buffer B before the next transform writes A then B
            }
        }
        f64w2_inverse_acc(t, s0, bufA, bufB, ti2, ti3, tut, bar_id, acc);
        bar_sync(bar_id, 64);
        f64w2_inverse_acc(t, s1, bufA, bufB, ti2, ti3, tut, bar_id, acc + 1024);
        bar_sync(bar_id, 64);   // acc is complete before the next step's rotated reads (other threads' words)
    }

    // ---- epilogue: sample_extract_index(0) (trlwe.rs:110-121) + key-switch digits (tlwe.rs:47-64) ----
    if (a.trlwe_out) {
        uint32_t* dst = a.trlwe_out + (size_t)gate * 2048;
        for (int k = t; k < 2048; k += 64) dst[k] = acc[k];
    }
    if (a.ksdig || a.lwe1_out) {
        for (int i = t; i < 1024; i += 64) {
            const uint32_t ai = (i == 0) ? acc[1024] : 0u - acc[1024 + 1024 - i];
            if (a.ksdig) a.ksdig[(size_t)gate * 1024 + i] = (uint16_t)((ai + 0x8000u) >> 16);
            if (a.lwe1_out) a.lwe1_out[(size_t)gate * 1025 + 1 + i] = ai;
        }
        if (a.lwe1_out && t == 0) a.lwe1_out[(size_t)gate * 1025] = acc[0];
    }
    if (a.out_init) {
        uint32_t* dst = a.out_init + (size_t)(a.idxo ? (long)a.idxo[gate] : gate) * (LWE_N + 1);
        for (int c = t; c <= LWE_N; c += 64) dst[c] = (c == 0) ? acc[0] : 0u;
    }
}
