// fft64.cuh -- the polynomial product of the FFT64 arithmetic mode: negacyclic f64 transform over C with EXACT integer rounding.
//
// The reference multiplies polynomials with an f64 FFT (utils/src/spqlios/fft_processor_spqlios.cpp:58-183 driving
// spqlios-{fft,ifft}-avx.s; pointwise stage utils/src/spqlios.rs:204-222; call sites hom_nand/src/trgsw.rs:264-306).  On the
// B200 a DFMA issues at the rate of an IMAD (64 per clock and SM, profiles/intpipe_r01b.json), so one complex f64 transform of
// 512 points (13.8 k DFMA-class operations) replaces a 1024-point NTT over a 29-bit prime (20.5 k IMAD slots), and -- the
// larger gain -- a 53-bit mantissa carries the whole 32-bit key at once: 6 forward + 2 inverse transforms and 12 pointwise
// polynomial products per CMUX instead of 6 + 4 and 24 (two 16-bit key slices) or 6 + 6 and 36 (three 11-bit slices).
//
// Exactness.  A coefficient of the external product is an integer of magnitude < 6144 * 32 * 2^31 < 2^49.  The transform
// computes it with an absolute rounding error far below 1/2 (measured on honestly generated keys: < 2^-8),
// so adding 1.5 * 2^52 leaves the EXACT integer mod 2^32 in the low word of the sum: the results are the same bits as the exact
// NTT modes', which the parity tests check.  Not a worst-case guarantee (neither is the reference's): the NTT modes stay
// selectable (tfhe_b200_set_key_slices).
//
// Mapping: a real polynomial a is folded into z_j = a_j + i a_{j+512}, j < 512, and evaluated at psi^(4k+1), psi = exp(i pi/1024).
// One WARP owns one polynomial, 16 complex values per lane:
//   forward  : pass A = stages 0..3 on the register index r (j = 32 r + lane; warp-uniform twiddles, immediates)
//              transpose through 8 KB of shared memory (16-byte elements, XOR swizzle, conflict free both ways)
//              pass B = stages 4..7 on rho = j[4:1] (lane' = 2 j[8:5] + j[0]; per-lane twiddles, 8 LDS.128)
//              stage 8 = lane pair exchange (32 SHFL): each lane of a pair takes 8 of the 16 butterflies
//              -> lane = p >> 4, register = p & 15 of position p, which holds the value at psi^(1 + 4 bitrev9(p))
//   inverse  : textbook decimation in time on that order: stages 0..3 in registers (constants, mostly trivial), stage 4 = lane
//              pair exchange, transpose, stages 5..8 on r (per-lane twiddles), z_j = psi^-j v_j, round, accumulate.
// Twiddles of sibling nodes differ by a factor i, which costs nothing in a fused butterfly: 12 per-lane loads forward, 8 inverse.
// 1/512 is folded into the transformed key.  Every function is what ONE lane does between two warp-level synchronisation
// points; all arithmetic goes through F_ADD / F_MUL / F_FMA (single IEEE operations, never contracted differently), so the
// CPU emulation (host_emul.cpp, built with -ffp-contract=off) produces the same bits as the GPU.
#pragma once
#include <stdint.h>
#include <string.h>
#include <utility>
#include "fft64_tables.h"

#if defined(__CUDACC__)
#define TFHE_HD __host__ __device__ __forceinline__
#else
#define TFHE_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define F_ADD(a, b) __dadd_rn((a), (b))
#define F_MUL(a, b) __dmul_rn((a), (b))
#define F_FMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define F_ADD(a, b) ((a) + (b))
#define F_MUL(a, b) ((a) * (b))
#define F_FMA(a, b, c) __builtin_fma((a), (b), (c))
#endif

namespace tfhe {

struct cd { double re, im; };
struct alignas(16) cd16 { double re, im; };   // what the tables, the transposes and the key hold (one LDS.128 / LDG.128)

constexpr int F64_PTS = 512;                 // complex points per polynomial
constexpr int F64_PER_LANE = 16;
constexpr int F64_FWDB_ROWS = 9, F64_INVA_ROWS = 8, F64_UNTW_ROWS = 16;
constexpr int F64_TAB_ELEMS = (F64_FWDB_ROWS + F64_INVA_ROWS) * 32;   // per-lane twiddle tables, [row][lane] cd16 (+ the untwist table)
constexpr double F64_ROUND_MAGIC = 6755399441055744.0;                             // 1.5 * 2^52
constexpr double F64_DIGIT_BIAS = 4503599627370496.0 + 32.0;                       // 2^52 + 32
// device key: [step i][row j < 6][output poly o < 2][register < 16][lane < 32] cd16 = 96 KB per step, 16 B per complex point
constexpr size_t F64_CHUNK_ELEMS = 512;                                            // one (row, output) polynomial: 8 KB
constexpr size_t F64_STEP_ELEMS = 12 * F64_CHUNK_ELEMS;
TFHE_HD size_t f64_key_off(int i, int j, int o) { return ((size_t)i * 12 + (size_t)j * 2 + o) * F64_CHUNK_ELEMS; }

static const double h_f64_fwdB[F64_FWDB_ROWS * 32 * 2] = {FFT64_FWD_B_LIST};
static const double h_f64_invA[F64_INVA_ROWS * 32 * 2] = {FFT64_INV_A_LIST};
static const double h_f64_untw[F64_UNTW_ROWS * 32 * 2] = {FFT64_UNTWIST_LIST};
#if defined(__CUDACC__)
static __device__ const double g_f64_fwdB[F64_FWDB_ROWS * 32 * 2] = {FFT64_FWD_B_LIST};
static __device__ const double g_f64_invA[F64_INVA_ROWS * 32 * 2] = {FFT64_INV_A_LIST};
static __device__ const double g_f64_untw[F64_UNTW_ROWS * 32 * 2] = {FFT64_UNTWIST_LIST};
#endif

// warp-uniform twiddles: on the device they sit in the constant bank and are read as DFMA operands (c[bank][offset]); as
// literals every one of them costs two 32-bit moves, a 64-bit immediate does not exist
static const double h_f64_fwdA[30] = {FFT64_FWD_A_LIST};
static const double h_f64_invC[40] = {FFT64_INV_C_LIST};
#if defined(__CUDACC__)
static __constant__ double c_f64_fwdA[30] = {FFT64_FWD_A_LIST};
static __constant__ double c_f64_invC[40] = {FFT64_INV_C_LIST};
#endif
template <int K> TFHE_HD double fwdA_c() {
#if defined(__CUDA_ARCH__)
    return c_f64_fwdA[K];
#else
    return h_f64_fwdA[K];
#endif
}
template <int K> TFHE_HD double invC_c() {
#if defined(__CUDA_ARCH__)
    return c_f64_invC[K];
#else
    return h_f64_invC[K];
#endif
}

TFHE_HD double f64_from_words(uint32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
    return __hiloint2double((int)hi, (int)lo);
#else
    const uint64_t b = ((uint64_t)hi << 32) | lo;
    double d;
    memcpy(&d, &b, 8);
    return d;
#endif
}
TFHE_HD uint32_t f64_low_word(double d) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(d);
#else
    uint64_t b;
    memcpy(&b, &d, 8);
    return (uint32_t)b;
#endif
}
TFHE_HD double f64_neg(double d) { return -d; }

// ---- butterflies: (a, b) <- (a + w b, a - w b), 6 operations; the second output is 2a - (a + w b) ----
TFHE_HD void bf_w(cd& a, cd& b, double wr, double wi) {
    const double pr = F_FMA(wr, b.re, F_FMA(-wi, b.im, a.re));
    const double pi = F_FMA(wr, b.im, F_FMA(wi, b.re, a.im));
    b.re = F_FMA(2.0, a.re, -pr); b.im = F_FMA(2.0, a.im, -pi);
    a.re = pr; a.im = pi;
}
// twiddle i w:  i w b = (-(wr b.im + wi b.re), wr b.re - wi b.im)
TFHE_HD void bf_iw(cd& a, cd& b, double wr, double wi) {
    const double pr = F_FMA(-wr, b.im, F_FMA(-wi, b.re, a.re));
    const double pi = F_FMA(wr, b.re, F_FMA(-wi, b.im, a.im));
    b.re = F_FMA(2.0, a.re, -pr); b.im = F_FMA(2.0, a.im, -pi);
    a.re = pr; a.im = pi;
}
// twiddle -i w:  -i w b = (wr b.im + wi b.re, -(wr b.re - wi b.im))
TFHE_HD void bf_miw(cd& a, cd& b, double wr, double wi) {
    const double pr = F_FMA(wr, b.im, F_FMA(wi, b.re, a.re));
    const double pi = F_FMA(-wr, b.re, F_FMA(wi, b.im, a.im));
    b.re = F_FMA(2.0, a.re, -pr); b.im = F_FMA(2.0, a.im, -pi);
    a.re = pr; a.im = pi;
}
TFHE_HD void bf_1(cd& a, cd& b) {   // w = 1
    const double sr = F_ADD(a.re, b.re), si = F_ADD(a.im, b.im);
    b.re = F_ADD(a.re, -b.re); b.im = F_ADD(a.im, -b.im);
    a.re = sr; a.im = si;
}
TFHE_HD void bf_mi(cd& a, cd& b) {  // w = -i: w b = (b.im, -b.re)
    const double sr = F_ADD(a.re, b.im), si = F_ADD(a.im, -b.re);
    const double dr = F_ADD(a.re, -b.im), di = F_ADD(a.im, b.re);
    a.re = sr; a.im = si; b.re = dr; b.im = di;
}
// sel: 0 = w, 1 = i w, 2 = -i w
template <int SEL>
TFHE_HD void bf_sel(cd& a, cd& b, double wr, double wi) {
    if (SEL == 0) bf_w(a, b, wr, wi);
    else if (SEL == 1) bf_iw(a, b, wr, wi);
    else bf_miw(a, b, wr, wi);
}

// =====================================================================================================
// forward
// =====================================================================================================
// pass A, stage S (0..3), butterfly I (0..7): registers r = 2 h beta + t and r + h, h = 8 >> S, warp-uniform twiddle
template <int S, int I>
TFHE_HD void f64_fa_bfly(cd (&x)[16]) {
    constexpr int h = 8 >> S, beta = I / h, t = I % h, ia = 2 * h * beta + t, k = (1 << S) - 1 + beta;
    bf_w(x[ia], x[ia + h], fwdA_c<2 * k>(), fwdA_c<2 * k + 1>());
}
template <int S, int... I>
TFHE_HD void f64_fa_stage(cd (&x)[16], std::integer_sequence<int, I...>) { (f64_fa_bfly<S, I>(x), ...); }
TFHE_HD void f64_fwd_passA(cd (&x)[16]) {
    f64_fa_stage<0>(x, std::make_integer_sequence<int, 8>{});
    f64_fa_stage<1>(x, std::make_integer_sequence<int, 8>{});
    f64_fa_stage<2>(x, std::make_integer_sequence<int, 8>{});
    f64_fa_stage<3>(x, std::make_integer_sequence<int, 8>{});
}
// transpose 1: element j = 32 r + lane is stored at slot j ^ (r & 7).  Lane' = 2 hi + top then needs BOTH halves of the block
// hi = j >> 5 (elements (hi, top' = 0 / 1, q), q = j & 15) for stage 4, whose butterflies pair top' = 0 with top' = 1: it reads
// the 32 elements and computes only ITS output of every butterfly (f64_t1_load_cross) -- the two lanes of a pair read the same
// addresses (16 distinct 16-byte words per instruction, two per bank group: two wavefronts).  A 16-values-per-lane transform of
// 512 points has one stage that pairs values of two lanes; taken here, on the load side of the transpose, it costs 16 loads and
// 16 operations more than half of its butterflies, instead of 32 shuffles and 96 selects after pass B.
TFHE_HD void f64_t1_store(int lane, const cd (&x)[16], cd16* S) {
    cd16* b[8];   // the swizzle only touches bits 0..2 of the lane: eight bases, everything else is an immediate offset
#pragma unroll
    for (int k = 0; k < 8; k++) b[k] = S + (lane ^ k);
#pragma unroll
    for (int r = 0; r < 16; r++) {
        cd16 v; v.re = x[r].re; v.im = x[r].im;
        b[r & 7][32 * r] = v;
    }
}
// stage 4 on the load side: x[q] = a_q + w b_q, w = +- twiddle of node (4, hi) (the sign is the lane's top bit, in the table)
TFHE_HD void f64_t1_load_cross(int lane, const cd16* S, const cd16& w, cd (&x)[16]) {
    const int hi = lane >> 1;
    const cd16* b[8];
#pragma unroll
    for (int k = 0; k < 8; k++) b[k] = S + ((hi << 5) | (k ^ (hi & 7)));
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int q = (k >> 1) | ((k & 1) << 3);   // 0, 8, 1, 9, ...: the operands of the first butterflies of pass B arrive first
        const cd16 A = b[q & 7][q & 8], B = b[q & 7][16 + (q & 8)];
        x[q].re = F_FMA(w.re, B.re, F_FMA(-w.im, B.im, A.re));
        x[q].im = F_FMA(w.re, B.im, F_FMA(w.im, B.re, A.im));
    }
}
// pass B: stages 5..8 on the register index q = j & 15, per-lane twiddles tb[t * 32 + lane] (row 0 is stage 4's).  The rows are
// loaded by f64_fwd_twB BEFORE the transpose (nothing can be hoisted above a warp barrier by the compiler), so their latency is
// under pass A and the transpose.
struct F64TwB { cd16 w[9]; };
TFHE_HD void f64_fwd_twB(int lane, const cd16* tb, F64TwB& t) {
#pragma unroll
    for (int k = 0; k < 9; k++) t.w[k] = tb[k * 32 + lane];
}
TFHE_HD void f64_fwd_passB(cd (&x)[16], const F64TwB& tw) {
    {
        const cd16 w = tw.w[1];
#pragma unroll
        for (int t = 0; t < 8; t++) bf_w(x[t], x[t + 8], w.re, w.im);
    }
    {
        const cd16 w = tw.w[2];
#pragma unroll
        for (int t = 0; t < 4; t++) { bf_w(x[t], x[t + 4], w.re, w.im); bf_iw(x[8 + t], x[12 + t], w.re, w.im); }
    }
#pragma unroll
    for (int c2 = 0; c2 < 2; c2++) {
        const cd16 w = tw.w[3 + c2];
#pragma unroll
        for (int t = 0; t < 2; t++) { bf_w(x[8 * c2 + t], x[8 * c2 + t + 2], w.re, w.im); bf_iw(x[8 * c2 + 4 + t], x[8 * c2 + 6 + t], w.re, w.im); }
    }
#pragma unroll
    for (int c2 = 0; c2 < 4; c2++) {
        const cd16 w = tw.w[5 + c2];
        bf_w(x[4 * c2], x[4 * c2 + 1], w.re, w.im);
        bf_iw(x[4 * c2 + 2], x[4 * c2 + 3], w.re, w.im);
    }
}
// the lane-pair exchange of the INVERSE transform's stage 4: lane parity j0 owns the butterflies m + 8 j0; it sends the operand
// it holds of the partner's butterflies
TFHE_HD void f64_x_send(int lane, const cd (&x)[16], cd (&send)[8]) {
    const bool odd = lane & 1;
#pragma unroll
    for (int m = 0; m < 8; m++) { send[m].re = odd ? x[m].re : x[m + 8].re; send[m].im = odd ? x[m].im : x[m + 8].im; }
}

// pointwise multiply-accumulate: acc[k] += y[k] * key[k * 32 + lane]; one half (registers 8 H .. 8 H + 7, 4 KB of key) at a time
template <int H>
TFHE_HD void f64_mac_half(int lane, const cd (&y)[16], const cd16* key /* the half's 256 values */, cd (&acc)[16]) {
#pragma unroll
    for (int k = 8 * H; k < 8 * H + 8; k++) {
        const cd16 w = key[(k - 8 * H) * 32 + lane];
        acc[k].re = F_FMA(y[k].re, w.re, F_FMA(-y[k].im, w.im, acc[k].re));
        acc[k].im = F_FMA(y[k].re, w.im, F_FMA(y[k].im, w.re, acc[k].im));
    }
}
TFHE_HD void f64_mac(int lane, const cd (&y)[16], const cd16* key, cd (&acc)[16]) {
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const cd16 w = key[k * 32 + lane];
        acc[k].re = F_FMA(y[k].re, w.re, F_FMA(-y[k].im, w.im, acc[k].re));
        acc[k].im = F_FMA(y[k].re, w.im, F_FMA(y[k].im, w.re, acc[k].im));
    }
}
TFHE_HD void f64_mul(int lane, const cd (&y)[16], const cd16* key, cd (&acc)[16]) {   // first row: acc = y * key
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const cd16 w = key[k * 32 + lane];
        acc[k].re = F_FMA(y[k].re, w.re, -F_MUL(y[k].im, w.im));
        acc[k].im = F_FMA(y[k].re, w.im, F_MUL(y[k].im, w.re));
    }
}

// =====================================================================================================
// inverse
// =====================================================================================================
template <int K> TFHE_HD void f64_inv_c(cd& a, cd& b) { bf_w(a, b, invC_c<2 * K>(), invC_c<2 * K + 1>()); }
// stages 0..3 on the register index (p & 15)
TFHE_HD void f64_inv_low(cd (&y)[16]) {
#pragma unroll
    for (int m = 0; m < 8; m++) bf_1(y[2 * m], y[2 * m + 1]);
#pragma unroll
    for (int c = 0; c < 4; c++) { bf_1(y[4 * c], y[4 * c + 2]); bf_mi(y[4 * c + 1], y[4 * c + 3]); }
#pragma unroll
    for (int c = 0; c < 2; c++) {
        bf_1(y[8 * c], y[8 * c + 4]);
        f64_inv_c<1>(y[8 * c + 1], y[8 * c + 5]);
        bf_mi(y[8 * c + 2], y[8 * c + 6]);
        f64_inv_c<3>(y[8 * c + 3], y[8 * c + 7]);
    }
    bf_1(y[0], y[8]);
    f64_inv_c<4 + 1>(y[1], y[9]); f64_inv_c<4 + 2>(y[2], y[10]); f64_inv_c<4 + 3>(y[3], y[11]);
    bf_mi(y[4], y[12]);
    f64_inv_c<4 + 5>(y[5], y[13]); f64_inv_c<4 + 6>(y[6], y[14]); f64_inv_c<4 + 7>(y[7], y[15]);
}
// stage 4 (span 16 = lane bit 0): lane parity lam owns the butterflies of registers m + 8 lam, m < 8; twiddle Wc^(16 m) (-i)^lam.
// v[m] = result with p[4] = 0, v[8 + m] = result with p[4] = 1; the lane now holds p[3] = lam.
template <int M> TFHE_HD void f64_inv_x_one(bool odd, const cd (&y)[16], const cd (&recv)[8], cd (&v)[16]) {
    cd a, b;
    a.re = odd ? recv[M].re : y[M].re; a.im = odd ? recv[M].im : y[M].im;
    // b0 = odd ? y[M + 8] : recv[M];  b = odd ? -i b0 : b0 = odd ? (b0.im, -b0.re) : b0
    b.re = odd ? y[M + 8].im : recv[M].re;
    b.im = odd ? f64_neg(y[M + 8].re) : recv[M].im;
    if (M == 0) bf_1(a, b); else f64_inv_c<12 + M>(a, b);
    v[M] = a; v[8 + M] = b;
}
template <int... M>
TFHE_HD void f64_inv_x_all(bool odd, const cd (&y)[16], const cd (&recv)[8], cd (&v)[16], std::integer_sequence<int, M...>) {
    (f64_inv_x_one<M>(odd, y, recv, v), ...);
}
TFHE_HD void f64_inv_x_bfly(int lane, const cd (&y)[16], const cd (&recv)[8], cd (&v)[16]) {
    f64_inv_x_all((lane & 1) != 0, y, recv, v, std::make_integer_sequence<int, 8>{});
}
// transpose 2: lane = 2 hi + lam holds p = (hi, p4, lam, m) in v[8 p4 + m]; slot = p ^ (((hi & 3) << 1) | lam);
// lane'' = p & 31 then reads p = 32 r + lane''
TFHE_HD void f64_t2_store(int lane, const cd (&v)[16], cd16* S) {
    const int hi = lane >> 1, lam = lane & 1;
    const int base = ((hi << 5) | (lam << 3)) ^ (((hi & 3) << 1) | lam);
    cd16* b[8];   // e & 7 meets the swizzle on bits 0..2 (eight bases); bit 4 is a plain offset
#pragma unroll
    for (int k = 0; k < 8; k++) b[k] = S + (base ^ k);
#pragma unroll
    for (int e = 0; e < 16; e++) {
        cd16 t; t.re = v[e].re; t.im = v[e].im;
        b[e & 7][(e >> 3) << 4] = t;
    }
}
TFHE_HD void f64_t2_load(int lane, const cd16* S, cd (&w)[16]) {
    const int l3 = (lane >> 3) & 1;
    const cd16* b[4];
#pragma unroll
    for (int k = 0; k < 4; k++) b[k] = S + ((lane ^ l3) ^ (k << 1));
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const cd16 t = b[r & 3][32 * r];
        w[r].re = t.re; w[r].im = t.im;
    }
}
// stages 5..8 on r, per-lane twiddles ta[t * 32 + lane] (loaded by f64_inv_twA before the transpose, like the forward rows)
struct F64TwA { cd16 w[8]; };
TFHE_HD void f64_inv_twA(int lane, const cd16* ta, F64TwA& t) {
#pragma unroll
    for (int k = 0; k < 8; k++) t.w[k] = ta[k * 32 + lane];
}
TFHE_HD void f64_inv_passA(cd (&w)[16], const F64TwA& tw) {
    {
        const cd16 t = tw.w[0];
#pragma unroll
        for (int c = 0; c < 8; c++) bf_w(w[2 * c], w[2 * c + 1], t.re, t.im);
    }
    {
        const cd16 t = tw.w[1];
#pragma unroll
        for (int c = 0; c < 4; c++) { bf_w(w[4 * c], w[4 * c + 2], t.re, t.im); bf_miw(w[4 * c + 1], w[4 * c + 3], t.re, t.im); }
    }
    {
        const cd16 t0 = tw.w[2], t1 = tw.w[3];
#pragma unroll
        for (int c = 0; c < 2; c++) {
            bf_w(w[8 * c], w[8 * c + 4], t0.re, t0.im);
            bf_w(w[8 * c + 1], w[8 * c + 5], t1.re, t1.im);
            bf_miw(w[8 * c + 2], w[8 * c + 6], t0.re, t0.im);
            bf_miw(w[8 * c + 3], w[8 * c + 7], t1.re, t1.im);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const cd16 t = tw.w[4 + k];
        bf_w(w[k], w[k + 8], t.re, t.im);
        bf_miw(w[k + 4], w[k + 12], t.re, t.im);
    }
}
TFHE_HD void f64_inv_passA(int lane, cd (&w)[16], const cd16* ta) {
    F64TwA tw;
    f64_inv_twA(lane, ta, tw);
    f64_inv_passA(w, tw);
}
// z_j = psi^-j v_j, exact rounding; lo[r] / hi[r] = coefficients j and j + 512 (j = 32 r + lane) mod 2^32.
// frac (host-side diagnostics only): largest distance of a value from the nearest integer.
TFHE_HD void f64_untwist_round(int lane, const cd (&w)[16], const cd16* ut, uint32_t (&lo)[16], uint32_t (&hi)[16], double* frac = nullptr) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const cd16 t = ut[r * 32 + lane];
        const double zr = F_FMA(w[r].re, t.re, -F_MUL(w[r].im, t.im));
        const double zi = F_FMA(w[r].re, t.im, F_MUL(w[r].im, t.re));
        const double mr = F_ADD(zr, F64_ROUND_MAGIC), mi = F_ADD(zi, F64_ROUND_MAGIC);
        lo[r] = f64_low_word(mr);
        hi[r] = f64_low_word(mi);
#if !defined(__CUDA_ARCH__)
        if (frac) {
            const double er = zr - (mr - F64_ROUND_MAGIC), ei = zi - (mi - F64_ROUND_MAGIC);
            const double m = (er < 0 ? -er : er) > (ei < 0 ? -ei : ei) ? (er < 0 ? -er : er) : (ei < 0 ? -ei : ei);
            if (m > *frac) *frac = m;
        }
#else
        (void)frac;
#endif
    }
}

// =====================================================================================================
// the same transform on TWO warps (latency kernel, blind_rotate_f64l2.cuh): 8 values per thread, three radix-8 passes
//   forward: pass 1 on j[8:6] (j = t + 64 e), pass 2 on j[5:3] (thread = 8 j[8:6] + j[2:0]), pass 3 on j[2:0] (thread = j >> 3)
//   inverse: pass 1' on p[2:0], pass 2' on p[5:3], pass 3' on p[8:6]; positions p as in the one-warp transform
// =====================================================================================================
static const double h_l2_fwd2[4 * 8 * 2] = {FFT64_L2_FWD2_LIST};
static const double h_l2_fwd3[4 * 64 * 2] = {FFT64_L2_FWD3_LIST};
static const double h_l2_inv2[4 * 8 * 2] = {FFT64_L2_INV2_LIST};
static const double h_l2_inv3[4 * 64 * 2] = {FFT64_L2_INV3_LIST};
static const double h_l2_untw[8 * 64 * 2] = {FFT64_L2_UNTWIST_LIST};
// three butterfly stages on 8 registers with warp-uniform twiddles: pass 1 of the forward transform (nodes (s, e >> (3 - s)))
template <int S, int I>
TFHE_HD void l2_f1_bfly(cd (&x)[8]) {
    constexpr int h = 4 >> S, beta = I / h, t = I % h, ia = 2 * h * beta + t, k = (1 << S) - 1 + beta;
    bf_w(x[ia], x[ia + h], fwdA_c<2 * k>(), fwdA_c<2 * k + 1>());
}
TFHE_HD void l2_fwd_pass1(cd (&x)[8]) {
    l2_f1_bfly<0, 0>(x); l2_f1_bfly<0, 1>(x); l2_f1_bfly<0, 2>(x); l2_f1_bfly<0, 3>(x);
    l2_f1_bfly<1, 0>(x); l2_f1_bfly<1, 1>(x); l2_f1_bfly<1, 2>(x); l2_f1_bfly<1, 3>(x);
    l2_f1_bfly<2, 0>(x); l2_f1_bfly<2, 1>(x); l2_f1_bfly<2, 2>(x); l2_f1_bfly<2, 3>(x);
}
// three stages with per-thread twiddles w[0] (first stage), w[1] (second; odd node: times i), w[2], w[3] (third; odd nodes: times i)
TFHE_HD void l2_fwd_pass23(cd (&x)[8], const cd16 (&w)[4]) {
#pragma unroll
    for (int m = 0; m < 4; m++) bf_w(x[m], x[m + 4], w[0].re, w[0].im);
    bf_w(x[0], x[2], w[1].re, w[1].im); bf_w(x[1], x[3], w[1].re, w[1].im);
    bf_iw(x[4], x[6], w[1].re, w[1].im); bf_iw(x[5], x[7], w[1].re, w[1].im);
    bf_w(x[0], x[1], w[2].re, w[2].im); bf_iw(x[2], x[3], w[2].re, w[2].im);
    bf_w(x[4], x[5], w[3].re, w[3].im); bf_iw(x[6], x[7], w[3].re, w[3].im);
}
// inverse pass 1': stages 0..2 on p[2:0], constants
TFHE_HD void l2_inv_pass1(cd (&y)[8]) {
    bf_1(y[0], y[1]); bf_1(y[2], y[3]); bf_1(y[4], y[5]); bf_1(y[6], y[7]);
    bf_1(y[0], y[2]); bf_mi(y[1], y[3]); bf_1(y[4], y[6]); bf_mi(y[5], y[7]);
    bf_1(y[0], y[4]); f64_inv_c<1>(y[1], y[5]); bf_mi(y[2], y[6]); f64_inv_c<3>(y[3], y[7]);
}
// inverse passes 2', 3': twiddles v[0] (first stage), v[1] (second; odd: times -i), v[2], v[3] (third: m = 0, 1; m = 2, 3: times -i)
TFHE_HD void l2_inv_pass23(cd (&y)[8], const cd16 (&v)[4]) {
#pragma unroll
    for (int c = 0; c < 4; c++) bf_w(y[2 * c], y[2 * c + 1], v[0].re, v[0].im);
    bf_w(y[0], y[2], v[1].re, v[1].im); bf_miw(y[1], y[3], v[1].re, v[1].im);
    bf_w(y[4], y[6], v[1].re, v[1].im); bf_miw(y[5], y[7], v[1].re, v[1].im);
    bf_w(y[0], y[4], v[2].re, v[2].im); bf_w(y[1], y[5], v[3].re, v[3].im);
    bf_miw(y[2], y[6], v[2].re, v[2].im); bf_miw(y[3], y[7], v[3].re, v[3].im);
}

// ---- inputs ----
// The masked source words of a polynomial are LANE-PRIVATE: lane l computes u[r] for the coefficients 32 r + l, r < 32, and the
// folded input of its transforms is x[r] = (digit(u[r]), digit(u[r + 16])), r < 16.  The three gadget digits of the 32 words are
// packed as bytes, one plane of 2 x 16 bytes per digit, and parked in shared memory between the transforms (32 live registers
// less): byte = top byte of (u << 6 dw) with its two low bits cleared = 4 * digit as a signed byte (the factor 4 is folded into
// the key scale).  A digit then costs ONE instruction (I2F.F64.S8 with a byte selector) and 1/16 of an LDS.128.
struct u4 { uint32_t x, y, z, w; };
constexpr double F64_KEY_SCALE = 1.0 / 2048.0;   // 1/512 of the inverse transform, 1/4 of the digit bytes
TFHE_HD uint32_t f64_byte_perm(uint32_t x, uint32_t y, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(x, y, s);
#else
    const uint64_t v = ((uint64_t)y << 32) | x;
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) r |= (uint32_t)((v >> (8 * ((s >> (4 * k)) & 7))) & 0xFF) << (8 * k);
    return r;
#endif
}
// plane word q (q < 8) of digit dw: bytes of u[4 q .. 4 q + 3]
template <int DW>
TFHE_HD uint32_t f64_pack4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    const uint32_t lo = f64_byte_perm(a << (6 * DW), b << (6 * DW), 0x0073u), hi = f64_byte_perm(c << (6 * DW), d << (6 * DW), 0x0073u);
    return f64_byte_perm(lo, hi, 0x5410u) & 0xFCFCFCFCu;
}
template <int DW>
TFHE_HD void f64_pack_plane(const uint32_t (&u)[32], u4& re, u4& im) {
    re.x = f64_pack4<DW>(u[0], u[1], u[2], u[3]);     re.y = f64_pack4<DW>(u[4], u[5], u[6], u[7]);
    re.z = f64_pack4<DW>(u[8], u[9], u[10], u[11]);   re.w = f64_pack4<DW>(u[12], u[13], u[14], u[15]);
    im.x = f64_pack4<DW>(u[16], u[17], u[18], u[19]); im.y = f64_pack4<DW>(u[20], u[21], u[22], u[23]);
    im.z = f64_pack4<DW>(u[24], u[25], u[26], u[27]); im.w = f64_pack4<DW>(u[28], u[29], u[30], u[31]);
}
TFHE_HD double f64_byte(uint32_t w, int k) { return (double)(int8_t)(w >> (8 * k)); }
TFHE_HD void f64_unpack4(uint32_t w, double& a, double& b, double& c, double& d) {
    a = f64_byte(w, 0); b = f64_byte(w, 1); c = f64_byte(w, 2); d = f64_byte(w, 3);
}
// folded input (scaled by 4) from one digit plane
TFHE_HD void f64_digits(const u4& re, const u4& im, cd (&x)[16]) {
    f64_unpack4(re.x, x[0].re, x[1].re, x[2].re, x[3].re);     f64_unpack4(re.y, x[4].re, x[5].re, x[6].re, x[7].re);
    f64_unpack4(re.z, x[8].re, x[9].re, x[10].re, x[11].re);   f64_unpack4(re.w, x[12].re, x[13].re, x[14].re, x[15].re);
    f64_unpack4(im.x, x[0].im, x[1].im, x[2].im, x[3].im);     f64_unpack4(im.y, x[4].im, x[5].im, x[6].im, x[7].im);
    f64_unpack4(im.z, x[8].im, x[9].im, x[10].im, x[11].im);   f64_unpack4(im.w, x[12].im, x[13].im, x[14].im, x[15].im);
}
// key polynomial (torus words taken as centred 32-bit integers) -> folded complex input
TFHE_HD void f64_key_input(int lane, const uint32_t* poly, cd (&x)[16]) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
        x[r].re = (double)(int32_t)poly[32 * r + lane];
        x[r].im = (double)(int32_t)poly[512 + 32 * r + lane];
    }
}

}  // namespace tfhe
