// ntt32.cuh -- arithmetic core of the B200 TFHE engine: exact negacyclic NTT over Z_p[X]/(X^1024+1).
//
// Replaces the reference's f64 spqlios FFT (utils/src/spqlios/spqlios-{fft,ifft}-avx.s driven by
// utils/src/spqlios/fft_processor_spqlios.cpp:58-183) with exact modular integer arithmetic.
//
// Design:
//  * one prime p < 2^29 (8p < 2^32): Shoup multiplication (2 IMAD + 1 IMAD.HI) with lazy Harvey ranges; the only
//    correction per butterfly is min(x, x-2p) which sm_100a executes as ONE instruction (VIADDMNMX.U32) on the ALU
//    pipe, off the FMA pipe that bounds the kernel (profiles/intpipe_r01.json).
//  * 1024 = 32 x 32: one WARP owns one polynomial, each thread holds 32 coefficients in registers and runs a fully
//    unrolled 32-point network twice (columns with warp-uniform twiddles read from the constant bank, rows with
//    per-thread twiddles read from shared memory), with a single swizzled shared-memory transpose in between.
//    No block-level barrier is needed inside a transform.
//  * forward = merged-twist Cooley-Tukey (natural in, bit-reversed out), inverse = Gentleman-Sande (mirror image);
//    the pointwise stage does not care about the order, 1/N and the Montgomery 2^32 are folded into the key.
//
// Everything here is __host__ __device__ so tests/host_emul can run the identical arithmetic on the CPU.
#pragma once
#include <stdint.h>
#include <utility>
#include "ntt_tables.h"

#if defined(__CUDACC__)
#define TFHE_HD __host__ __device__ __forceinline__
#else
#define TFHE_HD inline
#endif

namespace tfhe {

constexpr uint32_t P = NTT_P;
constexpr uint32_t P2 = NTT_2P;
constexpr int TWB_STRIDE = NTT_TWB_STRIDE;

static const uint32_t h_fwdA[64] = {NTT_FWD_A_LIST};
static const uint32_t h_invA[64] = {NTT_INV_A_LIST};
static const uint32_t h_fwdB[32 * NTT_TWB_STRIDE] = {NTT_FWD_B_LIST};
static const uint32_t h_invB[32 * NTT_TWB_STRIDE] = {NTT_INV_B_LIST};
#if defined(__CUDACC__)
static __constant__ uint32_t c_fwdA[64] = {NTT_FWD_A_LIST};
static __constant__ uint32_t c_invA[64] = {NTT_INV_A_LIST};
static __device__ const uint32_t g_fwdB[32 * NTT_TWB_STRIDE] = {NTT_FWD_B_LIST};
static __device__ const uint32_t g_invB[32 * NTT_TWB_STRIDE] = {NTT_INV_B_LIST};
static __constant__ uint32_t c_zero = 0;  // opaque to ptxas (a __constant__ may be rewritten by the host)
#endif

// Stage 0 of the forward COLUMN pass multiplies gadget digits d in [-32, 32) by the single twiddle psi^512: a 64-entry table
// IOTA[d + 32] = d * psi^512 mod p in [0, p) replaces those 16 Shoup multiplications per transform (DIGIT_TAB_WORDS words
// of shared memory behind the twiddle rows).
constexpr int DIGIT_TAB_WORDS = 64;
struct DigitTab { uint32_t v[DIGIT_TAB_WORDS]; };
TFHE_HD constexpr DigitTab make_digit_tab() {
    constexpr uint32_t fwdA[64] = {NTT_FWD_A_LIST};
    DigitTab t{};
    for (int k = 0; k < DIGIT_TAB_WORDS; k++) {
        const int d = k - 32;
        const uint64_t r = (uint64_t)(d < 0 ? (int64_t)NTT_P + d : d);
        t.v[k] = (uint32_t)(r * fwdA[2] % NTT_P);
    }
    return t;
}
static const DigitTab h_digit_tab = make_digit_tab();
// The same products for the throughput kernel (t2_steps.cuh), indexed by the RAW two's-complement 6-bit pattern of the digit
// (v = d & 63: no bias to add to the index) and pre-biased by p: IOTA2[v] = (d * psi^512 mod p) + p in [p, 2p).
TFHE_HD constexpr DigitTab make_digit_tab2() {
    constexpr uint32_t fwdA[64] = {NTT_FWD_A_LIST};
    DigitTab t{};
    for (int v = 0; v < DIGIT_TAB_WORDS; v++) {
        const int d = v < 32 ? v : v - 64;
        const uint64_t r = (uint64_t)(d < 0 ? (int64_t)NTT_P + d : d);
        t.v[v] = (uint32_t)(r * fwdA[2] % NTT_P) + NTT_P;
    }
    return t;
}
static const DigitTab h_digit_tab2 = make_digit_tab2();
#if defined(__CUDACC__)
static __device__ const DigitTab g_digit_tab = make_digit_tab();
static __device__ const DigitTab g_digit_tab2 = make_digit_tab2();
#endif

// a + b on the ALU pipe: max(a + b, 0) is ONE VIADDMNMX.U32, which ptxas cannot turn into IMAD.IADD -- plain additions
// stay off the FMA-heavy pipe that bounds the transforms (profiles/r01_ncu_blind_rotate_v1.txt).  (Round 1 used a three-input
// add with an opaque zero from the constant bank; ptxas hoisted and re-associated that zero into extra instructions.)
TFHE_HD uint32_t add_alu(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __viaddmax_u32(a, b, 0u);
#else
    return a + b;
#endif
}
// pins a value: the expression that produced it stays ONE instruction (a three-input IADD3) instead of being re-associated
// with its consumers -- ptxas otherwise splits a two-input add or subtract off it, often onto the FMA-heavy pipe (IMAD.IADD),
// and adds the constant again at every consumer.  Emits nothing.
TFHE_HD uint32_t pin(uint32_t v) {
#if defined(__CUDA_ARCH__)
    asm("" : "+r"(v));
#endif
    return v;
}
TFHE_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
// min(x, x - m) for unsigned x: conditional subtraction without a branch (VIADDMNMX.U32 on sm_100a)
TFHE_HD uint32_t csub(uint32_t x, uint32_t m) {
    uint32_t y = x - m;
#if defined(__CUDA_ARCH__)
    return min(x, y);
#else
    return y < x ? y : x;
#endif
}
// Shoup: y * w mod p for ANY 32-bit y, ws = floor(w * 2^32 / p); result in [0, 2p)
TFHE_HD uint32_t shoup_mul(uint32_t y, uint32_t w, uint32_t ws) { return y * w - mulhi32(y, ws) * P; }
// Montgomery reduction of S < p * 2^32 : returns S * 2^-32 mod p in (0, 2p)
TFHE_HD uint32_t redc64(uint64_t s) {
    uint32_t m = (uint32_t)s * NTT_PINV;
    return pin((uint32_t)(s >> 32) + P - mulhi32(m, P));
}
// centred lift of v in [0,p) to the exact signed integer (|true value| < p/2 by construction)
TFHE_HD int32_t lift(uint32_t v) { return v > (P - 1) / 2 ? (int32_t)(v - P) : (int32_t)v; }

// ---- twiddle providers: entry q in [1,32) = (w, shoup(w)) ----
template <bool INV>
struct TwUniform {  // pass A: same for every lane -> constant-bank operands
    TFHE_HD void get(int q, uint32_t& w, uint32_t& ws) const {
#if defined(__CUDA_ARCH__)
        w = INV ? c_invA[2 * q] : c_fwdA[2 * q];
        ws = INV ? c_invA[2 * q + 1] : c_fwdA[2 * q + 1];
#else
        w = INV ? h_invA[2 * q] : h_fwdA[2 * q];
        ws = INV ? h_invA[2 * q + 1] : h_fwdA[2 * q + 1];
#endif
    }
    TFHE_HD void get2(int q, uint32_t& w0, uint32_t& ws0, uint32_t& w1, uint32_t& ws1) const {
        get(q, w0, ws0);
        get(q + 1, w1, ws1);
    }
};
struct TwRow {  // pass B: per-thread row (shared memory on the device), 16-byte aligned
    const uint32_t* row;
    TFHE_HD void get(int q, uint32_t& w, uint32_t& ws) const {
        const uint2 v = *reinterpret_cast<const uint2*>(row + 2 * q);
        w = v.x;
        ws = v.y;
    }
    TFHE_HD void get2(int q, uint32_t& w0, uint32_t& ws0, uint32_t& w1, uint32_t& ws1) const {  // q even
        const uint4 v = *reinterpret_cast<const uint4*>(row + 2 * q);
        w0 = v.x; ws0 = v.y; w1 = v.z; ws1 = v.w;
    }
};

// ---- 32-point merged-twist Cooley-Tukey network on registers; values stay in [0,4p) ----
// CORR = 0: no correction (bound grows by 2p), CORR = 4: conditional subtraction of 4p first.  Any 32-bit value is a
// legal Shoup input, so only the pass-through operand needs its range watched; 8p < 2^32 leaves room for three
// uncorrected stages.  Range bookkeeping (upper bounds, in units of p) is in ct32 below.
template <int CORR>
TFHE_HD void ct_bfly(uint32_t& a, uint32_t& b, uint32_t w, uint32_t ws) {
    const uint32_t X = CORR ? csub(a, 2u * P2) : a;
    const uint32_t T = shoup_mul(b, w, ws);
    a = add_alu(X, T);
    b = pin(X - T + P2);
}
template <int S, int CORR, class TW>
TFHE_HD void ct_stage(uint32_t (&x)[32], const TW& tw) {  // stage S = 1..4: m = 2^S blocks of half-width t = 16 >> S
    constexpr int m = 1 << S, t = 16 >> S;
#pragma unroll
    for (int i = 0; i < m; i += 2) {
        uint32_t w0, ws0, w1, ws1;
        tw.get2(m + i, w0, ws0, w1, ws1);
#pragma unroll
        for (int j = 0; j < t; j++) ct_bfly<CORR>(x[2 * i * t + j], x[2 * i * t + j + t], w0, ws0);
#pragma unroll
        for (int j = 0; j < t; j++) ct_bfly<CORR>(x[2 * (i + 1) * t + j], x[2 * (i + 1) * t + j + t], w1, ws1);
    }
}
// Input bound x < 2p (ct32): bounds after each stage (a butterfly outputs < X + 2p, csub(.,4p) maps [0,8p) to [0,4p)):
//   s0: 4p   s1: 6p   s2: 8p   s3 (corrected to 4p): 6p   s4: 8p        -> output < 8p < 2^32
// Input bound x < 8p (ct32_wide, the row pass fed straight from the column pass):
//   s0 (corrected): 6p   s1: 8p   s2 (corrected): 6p   s3: 8p   s4 (corrected): 6p   -> output < 6p
template <int C0, int C1, int C2, int C3, int C4, class TW>
TFHE_HD void ct32_plan(uint32_t (&x)[32], const TW& tw) {
    {
        uint32_t w, ws;
        tw.get(1, w, ws);
#pragma unroll
        for (int j = 0; j < 16; j++) ct_bfly<C0>(x[j], x[j + 16], w, ws);
    }
    ct_stage<1, C1>(x, tw);
    ct_stage<2, C2>(x, tw);
    ct_stage<3, C3>(x, tw);
    ct_stage<4, C4>(x, tw);
}
template <class TW>
TFHE_HD void ct32(uint32_t (&x)[32], const TW& tw) { ct32_plan<0, 0, 0, 4, 0>(x, tw); }
// stages 1..4 only, for inputs < 2p whose stage 0 was done by table lookup (bounds 4p, 6p, corrected 6p, 8p)
template <class TW>
TFHE_HD void ct32_after_stage0(uint32_t (&x)[32], const TW& tw) {
    ct_stage<1, 0>(x, tw);
    ct_stage<2, 0>(x, tw);
    ct_stage<3, 4>(x, tw);
    ct_stage<4, 0>(x, tw);
}
template <class TW>
TFHE_HD void ct32_wide(uint32_t (&x)[32], const TW& tw) { ct32_plan<4, 0, 4, 0, 4>(x, tw); }
// ---- 32-point Gentleman-Sande network (exact mirror of ct32) with per-element lazy ranges ----
// A GS butterfly maps (U, V) to (U + V, (U - V + K p) w): the product output is always < 2p (Shoup takes any 32-bit
// input), the sum output has the SUM of the operand bounds.  Instead of one conditional subtraction per butterfly
// (80 per network) the bound of every element is tracked at compile time and an operand is halved only where
// bound(U) + bound(V) would pass 8p < 2^32: 24 conditional subtractions per network for inputs < 2p, none of them in
// stages 0 and 1.  K = bound(V) keeps the difference non-negative and below 2^32.
struct GsPlan {
    unsigned char cu[5][16][2], cv[5][16][2];   // conditional subtractions on U / V before butterfly b of stage s (units of p, 0 = none)
    unsigned char kv[5][16];                    // K of the difference
    unsigned char out[32];                      // bound of every output, units of p
};
TFHE_HD constexpr GsPlan make_gs_plan(int in) {
    GsPlan pl{};
    int bnd[32] = {};
    for (int i = 0; i < 32; i++) bnd[i] = in;
    for (int S = 0; S < 5; S++) {
        const int t = 1 << S, h = 16 >> S;
        for (int i = 0; i < h; i++)
            for (int j = 0; j < t; j++) {
                const int u = 2 * i * t + j, v = u + t, bi = i * t + j;
                int nu = 0, nv = 0;
                while (bnd[u] + bnd[v] > 8) {
                    if (bnd[u] >= bnd[v]) { const int m = (bnd[u] + 1) / 2; pl.cu[S][bi][nu++] = (unsigned char)m; bnd[u] = m; }
                    else { const int m = (bnd[v] + 1) / 2; pl.cv[S][bi][nv++] = (unsigned char)m; bnd[v] = m; }
                }
                pl.kv[S][bi] = (unsigned char)bnd[v];
                bnd[u] = bnd[u] + bnd[v];
                bnd[v] = 2;
            }
    }
    for (int i = 0; i < 32; i++) pl.out[i] = (unsigned char)bnd[i];
    return pl;
}
// butterfly BI (= block * t + j) of stage S of the plan for inputs < IN p
template <int IN, int S, int BI>
TFHE_HD void gs_bfly_p(uint32_t (&x)[32], uint32_t w, uint32_t ws) {
    constexpr GsPlan pl = make_gs_plan(IN);
    constexpr int t = 1 << S, i = BI / t, j = BI % t, u = 2 * i * t + j, v = u + t;
    uint32_t U = x[u], V = x[v];
    if constexpr (pl.cu[S][BI][0] != 0) U = csub(U, pl.cu[S][BI][0] * P);
    if constexpr (pl.cu[S][BI][1] != 0) U = csub(U, pl.cu[S][BI][1] * P);
    if constexpr (pl.cv[S][BI][0] != 0) V = csub(V, pl.cv[S][BI][0] * P);
    if constexpr (pl.cv[S][BI][1] != 0) V = csub(V, pl.cv[S][BI][1] * P);
    x[u] = add_alu(U, V);
    x[v] = shoup_mul(pin(U - V + pl.kv[S][BI] * P), w, ws);
}
template <int IN, int S, int I, class TW, int... J>   // blocks I and I+1 of stage S (16 >> S >= 2): one 16-byte twiddle load
TFHE_HD void gs_blockpair_p(uint32_t (&x)[32], const TW& tw, std::integer_sequence<int, J...>) {
    constexpr int t = 1 << S, h = 16 >> S;
    uint32_t w0, ws0, w1, ws1;
    tw.get2(h + I, w0, ws0, w1, ws1);
    (gs_bfly_p<IN, S, I * t + J>(x, w0, ws0), ...);
    (gs_bfly_p<IN, S, (I + 1) * t + J>(x, w1, ws1), ...);
}
template <int IN, int S, class TW, int... I2>
TFHE_HD void gs_stage_p(uint32_t (&x)[32], const TW& tw, std::integer_sequence<int, I2...>) {
    (gs_blockpair_p<IN, S, 2 * I2>(x, tw, std::make_integer_sequence<int, (1 << S)>{}), ...);
}
template <int IN, class TW, int... J>
TFHE_HD void gs_last_p(uint32_t (&x)[32], const TW& tw, std::integer_sequence<int, J...>) {
    uint32_t w, ws;
    tw.get(1, w, ws);
    (gs_bfly_p<IN, 4, J>(x, w, ws), ...);
}
template <int B, int T>
TFHE_HD void gs_norm_one(uint32_t& v) {   // [0, B p) -> [0, T p)
    if constexpr (B > T) {
        v = csub(v, ((B + 1) / 2) * P);
        gs_norm_one<(B + 1) / 2, T>(v);
    }
}
// brings every output of the IN-plan below T p (T = 2: input range of the next pass, 24 subtractions; T = 1: canonical, 56)
template <int IN, int T, int... C>
TFHE_HD void gs_norm_seq(uint32_t (&x)[32], std::integer_sequence<int, C...>) {
    constexpr GsPlan pl = make_gs_plan(IN);
    (gs_norm_one<pl.out[C], T>(x[C]), ...);
}
template <int T>
TFHE_HD void gs_norm(uint32_t (&x)[32]) { gs_norm_seq<2, T>(x, std::make_integer_sequence<int, 32>{}); }

// full network for inputs < 2p; outputs carry the plan's bounds (2p, 4p or 8p): follow with gs_norm<T>
template <class TW>
TFHE_HD void gs32_lazy(uint32_t (&x)[32], const TW& tw) {
    gs_stage_p<2, 0>(x, tw, std::make_integer_sequence<int, 8>{});
    gs_stage_p<2, 1>(x, tw, std::make_integer_sequence<int, 4>{});
    gs_stage_p<2, 2>(x, tw, std::make_integer_sequence<int, 2>{});
    gs_stage_p<2, 3>(x, tw, std::make_integer_sequence<int, 1>{});
    gs_last_p<2>(x, tw, std::make_integer_sequence<int, 16>{});
}
template <class TW>
TFHE_HD void gs32(uint32_t (&x)[32], const TW& tw) {   // values in [0,2p) in, [0,2p) out
    gs32_lazy(x, tw);
    gs_norm<2>(x);
}

// The network split for software pipelining: stages 0 and 1 only touch the four values of one 16-byte chunk, so the
// pointwise stage can run them chunk by chunk while it still waits for the next chunk's key and spectrum loads
// (gs32_head4), and the transform proper starts at stage 2 (gs32_tail).  gs32_lazy == 8 x gs32_head4 + gs32_tail.
TFHE_HD constexpr bool gs_head_is_uniform() {   // the plan for inputs < 2p has no subtraction in stages 0/1 and K = 2, (4,2)
    const GsPlan pl = make_gs_plan(2);
    for (int b = 0; b < 16; b++) {
        if (pl.cu[0][b][0] || pl.cv[0][b][0] || pl.cu[1][b][0] || pl.cv[1][b][0]) return false;
        if (pl.kv[0][b] != 2 || pl.kv[1][b] != ((b & 1) ? 2 : 4)) return false;
    }
    return true;
}
static_assert(gs_head_is_uniform(), "gs32_head4 hard-codes the lazy plan of stages 0 and 1");
TFHE_HD void gs_bfly_k(uint32_t& a, uint32_t& b, uint32_t kp, uint32_t w, uint32_t ws) {
    const uint32_t U = a, V = b;
    a = add_alu(U, V);
    b = shoup_mul(pin(U - V + kp), w, ws);
}
template <class TW>
TFHE_HD void gs32_head4(uint32_t& x0, uint32_t& x1, uint32_t& x2, uint32_t& x3, int q, const TW& tw) {
    uint32_t w0, ws0, w1, ws1, w2, ws2;
    tw.get2(16 + 2 * q, w0, ws0, w1, ws1);
    tw.get(8 + q, w2, ws2);
    gs_bfly_k(x0, x1, 2 * P, w0, ws0);
    gs_bfly_k(x2, x3, 2 * P, w1, ws1);
    gs_bfly_k(x0, x2, 4 * P, w2, ws2);
    gs_bfly_k(x1, x3, 2 * P, w2, ws2);
}
template <class TW>
TFHE_HD void gs32_tail(uint32_t (&x)[32], const TW& tw) {   // outputs carry the plan's bounds: follow with gs_norm<T>
    gs_stage_p<2, 2>(x, tw, std::make_integer_sequence<int, 2>{});
    gs_stage_p<2, 3>(x, tw, std::make_integer_sequence<int, 1>{});
    gs_last_p<2>(x, tw, std::make_integer_sequence<int, 16>{});
}

// position of element (row r, column c) of a 32x32 word tile in shared memory.  Rows are PADDED to 36 words: the scalar
// "lane = column, loop over rows" accesses stay consecutive, and the 128-bit "lane = row" accesses of a quarter warp
// land on banks 4(r+q) mod 32 -- all 32 banks, conflict free -- while every address is one base register plus an
// immediate (an XOR swizzle needs eight base registers, which the 80-register kernel kept recomputing on the FMA pipe).
constexpr int TILE_STRIDE = 36;
constexpr int TILE_WORDS = 32 * TILE_STRIDE;
TFHE_HD int swz(int r, int c) { return r * TILE_STRIDE + c; }
TFHE_HD int swz_chunk(int r, int q) { return r * TILE_STRIDE + 4 * q; }

// ---- decomposition pieces (reference: utils/src/math.rs:300-326 with the mask of math.rs:542-560) ----
// digit `dw` (0 = most significant) of x after the mask trick, sign-extended from 6 bits
TFHE_HD int32_t gadget_digit(uint32_t x, uint32_t mask, int dw) {
    const uint32_t u = add_alu(x, mask) ^ mask;
    return ((int32_t)(u << (6 * dw))) >> 26;
}
// residue of a small signed integer: for v < 0 the unsigned value is huge and v + p wraps to the residue, so it is
// min(v, v + p) in unsigned arithmetic -- one VIADDMNMX
TFHE_HD uint32_t to_residue(int32_t v) { return csub((uint32_t)v, 0u - P); }

// centred slices of a torus word.
//   ns = 3 (default, exact in the worst case): 11/11/10 bits, c0 + 2^11 c1 + 2^22 c2 == C (mod 2^32), |c0|,|c1| <= 1024, |c2| <= 512
//   ns = 2 (opt-in fast mode):  16/16 bits,    c0 + 2^16 c1 == C (mod 2^32),           |c0|,|c1| <= 32768
TFHE_HD int slice_shift(int ns) { return ns == 3 ? 11 : 16; }
TFHE_HD int32_t key_slice(uint32_t C, int part, int ns = 3) {
    if (ns == 2) {
        const int32_t c0 = (int32_t)((C & 0xFFFFu) ^ 0x8000u) - 0x8000;
        if (part == 0) return c0;
        const uint32_t C1 = ((C - (uint32_t)c0) >> 16) & 0xFFFFu;
        return (int32_t)(C1 ^ 0x8000u) - 0x8000;
    }
    const int32_t c0 = (int32_t)((C & 0x7FFu) ^ 0x400u) - 0x400;
    if (part == 0) return c0;
    const uint32_t C1 = (C - (uint32_t)c0) >> 11;
    const int32_t c1 = (int32_t)((C1 & 0x7FFu) ^ 0x400u) - 0x400;
    if (part == 1) return c1;
    const uint32_t C2 = ((C1 - (uint32_t)c1) >> 11) & 0x3FFu;
    return (int32_t)((C2 & 0x3FFu) ^ 0x200u) - 0x200;
}

}  // namespace tfhe
