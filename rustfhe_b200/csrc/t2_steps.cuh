// t2_steps.cuh -- per-lane pieces of one CMUX step of the THROUGHPUT blind rotation (blind_rotate_t2.cuh): one gate on TWO
// warps, two 16-bit key slices.
//
// Reference semantics: hom_nand/src/tfhe.rs:103-110 (one fold step of blind_rotate), hom_nand/src/trgsw.rs:264-306
// (TRGSWRepF::cross: 6 iFFT + 12 hadamard-accumulate + 2 FFT), utils/src/math.rs:85-113 (rotate), math.rs:300-326 with the
// mask of math.rs:542-560 (decomposition_i32_).
//
// Work split.  Warp pw of a gate owns polynomial pw of the accumulator (0 = b / `cipher`, 1 = a / `p_key`):
//   phase 1 : u = ((X^abar acc - acc)[pw] + mask) ^ mask stays in REGISTERS (32 words per lane); the warp forward-transforms
//             its three gadget digits into the spectrum tiles dh[3 pw + 0..2]                   (b digits first, F10)
//   barrier : the six spectra of the gate are complete
//   phase 2 : the warp multiply-accumulates BOTH 16-bit key slices of output polynomial pw against the six spectra (every
//             spectrum chunk is read once for the two slices), and runs the inverse row passes
//   barrier : both warps have finished reading the spectra -> the own tiles are free as transpose scratch
//   phase 3 : inverse column passes, exact lift, acc[pw] += x0 + (x1 << 16)  (the warp is the only writer of acc[pw])
// Per CMUX: 6 forward + 4 inverse transforms and 24 pointwise polynomial products, against 6 + 6 and 36 with three 11-bit
// slices.  Spectra are left UNNORMALISED (< 8p + 32): the 64-bit pointwise sums (6 * 8p * p = 48 p^2 < 2^64) and one Montgomery
// reduction absorb the range, the inverse row pass starts from a < 8p plan (ntt32.cuh, GsPlan).
//
// Tiles are 32 x 32 words, XOR-swizzled on the 16-byte chunk index (no padding: six gates of 33 KB fit one SM).
// Every function is what ONE lane does between two warp-level synchronisation points (host_emul.cpp runs them lane by lane).
#pragma once
#include "cmux_steps.cuh"

namespace tfhe {

constexpr int T2_TILE_WORDS = 1024;
// element (row r, column c): the 16-byte chunk c/4 of row r sits at chunk position (c/4) ^ (r & 7).  Scalar column accesses
// (lane = c, fixed r) cover the 32 words of a row: conflict free; 128-bit row accesses (lane = r) of a quarter warp hit eight
// different chunk positions: conflict free.
TFHE_HD int xs(int r, int c) { return r * 32 + ((((c >> 2) ^ (r & 7)) << 2) | (c & 3)); }
TFHE_HD int xs_chunk(int r, int q) { return r * 32 + ((q ^ (r & 7)) << 2); }

// device BK layout of the throughput kernel: [step i][poly pw][chunk q][row j][slice s][lane][4] words, values in [0,p),
// pre-multiplied by 2^32 / N.  What one warp streams per step (48 KB) is contiguous, in the order it consumes it.
constexpr size_t T2_STEP_WORDS = (size_t)2 * 8 * BK_ROWS * 2 * 32 * 4;   // 24576 words = 96 KB = 8 B per transform coefficient
constexpr size_t T2_Q_STRIDE = (size_t)BK_ROWS * 2 * 32 * 4;             // words between two chunks of one (row, slice)
TFHE_HD size_t t2_bk_off(int i, int pw, int q, int j, int s, int lane) {
    return ((((((size_t)i * 2 + pw) * 8 + q) * BK_ROWS + j) * 2 + s) * 32 + lane) * 4;
}

// ---- phase 1u: lane = column c; u[r] = masked source coefficient 32 r + c ----
// u = ((X^abar A - A)[k] + mask) ^ mask with k = 32 r + lane (math.rs:85-113: X^abar A [k] = +-A[(k - abar) mod N], the sign
// flips once where the index folds around and once more for abar >= N).  Written in byte offsets with explicit sign masks so
// that one element costs 7 ALU instructions and two LDS: the ALU pipe is as loaded as the FMA pipe in this kernel.
// FLIP (abar >= N) is warp-uniform: the caller branches once per step and the NOT of the sign mask folds into the LOP3 / the
// constant of each version.
template <bool FLIP>
TFHE_HD void t2_u_rot(int lane, const uint32_t* A, uint32_t ap /* abar & 1023 */, uint32_t mask, uint32_t (&u)[32]) {
    const uint32_t tb = ((uint32_t)lane - ap) * 4u;                 // byte offset of coefficient k - ap for r = 0 (may wrap below zero)
    const uint32_t cm = FLIP ? mask + 1u : mask;                    // -(~m) = m + 1: the +1 rides on the mask addition
#pragma unroll
    for (int r = 0; r < 32; r++) {
        const uint32_t t4 = add_alu(tb, 128u * r);
        const uint32_t m = (uint32_t)((int32_t)t4 >> 31);           // ~0 where the rotated index folded around
        const uint32_t rv = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(A) + (t4 & 4092u));
        const uint32_t a = A[32 * r + lane];
        // FLIP = false: (rv ^ m) - m - a ;  FLIP = true: (rv ^ ~m) - ~m - a = ~(rv ^ m) + m + 1 - a
        const uint32_t d = pin(FLIP ? (~(rv ^ m)) + m - a : (rv ^ m) - m - a);
        u[r] = add_alu(d, cm) ^ mask;
    }
}
template <bool ROTATE>
TFHE_HD void t2_u(int lane, const uint32_t* A, uint32_t abar, uint32_t mask, uint32_t (&u)[32]) {
    if (!ROTATE) {
#pragma unroll
        for (int r = 0; r < 32; r++) u[r] = add_alu(A[32 * r + lane], mask) ^ mask;
    } else if (abar >> 10) {
        t2_u_rot<true>(lane, A, abar & 1023u, mask, u);
    } else {
        t2_u_rot<false>(lane, A, abar & 1023u, mask, u);
    }
}
// ---- phase 1a: digit dw of the column, column pass, scatter into tile S ----
// Stage 0 pairs rows r and r+16 with the single twiddle psi^512.  With X = d_r (a small SIGNED digit, taken straight out of
// the masked word by an arithmetic shift) and T' = (d_{r+16} psi^512 mod p) + p from the 64-entry table indexed by the raw
// 6-bit pattern:  a = X + T'  in [p-32, 2p+32),  b = X - T' + 3p  in (p-32, 2p+32]: no residue conversion, no index bias.
// The +-32 of slack is carried through the lazy bounds (8p + 32 < 2^32).  Corrections: last column stage only.
TFHE_HD void t2_fwd_cols(int lane, const uint32_t (&u)[32], int sh /* 6 * dw */, uint32_t* S, const uint32_t* digit_tab2) {
    uint32_t x[32];
    const int sh2 = 24 - sh;
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const uint32_t X = (uint32_t)(((int32_t)(u[r] << sh)) >> 26);
        const uint32_t T = *reinterpret_cast<const uint32_t*>(reinterpret_cast<const char*>(digit_tab2) + ((u[r + 16] >> sh2) & 0xFCu));
        x[r] = X + T;
        x[r + 16] = pin(X - T + 3u * P);
    }
    // bounds: in < 2p+32; s1 4p, s2 6p, s3 8p (+32); s4 corrected (< 4p+32) -> < 6p+32
    ct_stage<1, 0>(x, TwUniform<false>());
    ct_stage<2, 0>(x, TwUniform<false>());
    ct_stage<3, 0>(x, TwUniform<false>());
    ct_stage<4, 4>(x, TwUniform<false>());
#pragma unroll
    for (int r = 0; r < 32; r++) S[xs(r, lane)] = x[r];
}
// ---- phase 1b: lane = row.  Row pass, spectrum left in [0, 8p + 32), stored back in row layout ----
template <class TW>
TFHE_HD void t2_fwd_rows(int lane, uint32_t* S, const TW& tw) {
    uint32_t x[32];
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint4 v = *reinterpret_cast<const uint4*>(S + xs_chunk(lane, q));
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
    ct32_plan<0, 4, 0, 4, 0>(x, tw);   // column pass leaves < 6p+32: 8p, corrected 6p, 8p, corrected 6p, 8p (+32) -> < 8p + 32
#pragma unroll
    for (int q = 0; q < 8; q++)
        *reinterpret_cast<uint4*>(S + xs_chunk(lane, q)) = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
}

// ---- inverse networks for inputs < IN p (the plan machinery of ntt32.cuh with the input bound as a parameter) ----
template <int IN, int Q, class TW>
TFHE_HD void gs_head_q(uint32_t (&x)[32], const TW& tw) {   // stages 0 and 1 on the four values of chunk Q
    uint32_t w0, ws0, w1, ws1, w2, ws2;
    tw.get2(16 + 2 * Q, w0, ws0, w1, ws1);
    tw.get(8 + Q, w2, ws2);
    gs_bfly_p<IN, 0, 2 * Q>(x, w0, ws0);
    gs_bfly_p<IN, 0, 2 * Q + 1>(x, w1, ws1);
    gs_bfly_p<IN, 1, 2 * Q>(x, w2, ws2);
    gs_bfly_p<IN, 1, 2 * Q + 1>(x, w2, ws2);
}
template <int IN, class TW>
TFHE_HD void gs_tail_in(uint32_t (&x)[32], const TW& tw) {
    gs_stage_p<IN, 2>(x, tw, std::make_integer_sequence<int, 2>{});
    gs_stage_p<IN, 3>(x, tw, std::make_integer_sequence<int, 1>{});
    gs_last_p<IN>(x, tw, std::make_integer_sequence<int, 16>{});
}
template <int IN, int T>
TFHE_HD void gs_norm_in(uint32_t (&x)[32]) { gs_norm_seq<IN, T>(x, std::make_integer_sequence<int, 32>{}); }

constexpr int T2_MAC_IN = 8;   // pointwise sums of six (< 8p+32) x (< p) products, Montgomery-reduced: < 48 p^2 / 2^32 + p < 7.001 p < 2^32

// Montgomery reduction of S < 2^64 with S / 2^32 + p < 2^32: S * 2^-32 mod p, in (S/2^32 - p, S/2^32 + p]
TFHE_HD uint32_t redc64_wide(uint64_t s) { return redc64(s); }

// ---- phase 2: lane = row.  Both key slices of output poly pw against the six spectra; inverse row pass ----
// key: this warp's 48 KB of the step, [q][j][s][lane][4].  dh: six spectrum tiles.  y0 / y1: row `lane` of slice 0 / 1 after
// the inverse ROW pass, values < 2p.
template <int Q, class TW>
TFHE_HD void t2_mac_chunk(int lane, const uint32_t* key, const uint32_t* dh, const TW& tw, uint32_t (&y0)[32], uint32_t (&y1)[32]) {
    uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0, b0 = 0, b1 = 0, b2 = 0, b3 = 0;
#pragma unroll
    for (int j = 0; j < BK_ROWS; j++) {
        const uint4 d = *reinterpret_cast<const uint4*>(dh + j * T2_TILE_WORDS + xs_chunk(lane, Q));
#if defined(T2_EXP_NOKEY)   // timing experiment only: no key traffic
        const uint4 k0 = make_uint4(lane + Q, j + 7u, d.y, d.w), k1 = make_uint4(d.z, lane * 3u, Q + 1u, d.x);
#elif defined(__CUDA_ARCH__)
        const uint4 k0 = __ldg(reinterpret_cast<const uint4*>(key) + ((Q * BK_ROWS + j) * 2 + 0) * 32 + lane);
        const uint4 k1 = __ldg(reinterpret_cast<const uint4*>(key) + ((Q * BK_ROWS + j) * 2 + 1) * 32 + lane);
#else
        const uint4 k0 = *(reinterpret_cast<const uint4*>(key) + ((Q * BK_ROWS + j) * 2 + 0) * 32 + lane);
        const uint4 k1 = *(reinterpret_cast<const uint4*>(key) + ((Q * BK_ROWS + j) * 2 + 1) * 32 + lane);
#endif
        a0 += (uint64_t)d.x * k0.x; a1 += (uint64_t)d.y * k0.y; a2 += (uint64_t)d.z * k0.z; a3 += (uint64_t)d.w * k0.w;
        b0 += (uint64_t)d.x * k1.x; b1 += (uint64_t)d.y * k1.y; b2 += (uint64_t)d.z * k1.z; b3 += (uint64_t)d.w * k1.w;
    }
    y0[4 * Q] = redc64_wide(a0); y0[4 * Q + 1] = redc64_wide(a1); y0[4 * Q + 2] = redc64_wide(a2); y0[4 * Q + 3] = redc64_wide(a3);
    y1[4 * Q] = redc64_wide(b0); y1[4 * Q + 1] = redc64_wide(b1); y1[4 * Q + 2] = redc64_wide(b2); y1[4 * Q + 3] = redc64_wide(b3);
    gs_head_q<T2_MAC_IN, Q>(y0, tw);
    gs_head_q<T2_MAC_IN, Q>(y1, tw);
}
template <class TW, int... Q>
TFHE_HD void t2_mac_chunks(int lane, const uint32_t* key, const uint32_t* dh, const TW& tw, uint32_t (&y0)[32], uint32_t (&y1)[32],
                           std::integer_sequence<int, Q...>) {
    (t2_mac_chunk<Q>(lane, key, dh, tw, y0, y1), ...);
}
template <class TW>
TFHE_HD void t2_mac(int lane, const uint32_t* key, const uint32_t* dh, const TW& tw, uint32_t (&y0)[32], uint32_t (&y1)[32]) {
    t2_mac_chunks(lane, key, dh, tw, y0, y1, std::make_integer_sequence<int, 8>{});
    gs_tail_in<T2_MAC_IN>(y0, tw);
    gs_norm_in<T2_MAC_IN, 2>(y0);
    gs_tail_in<T2_MAC_IN>(y1, tw);
    gs_norm_in<T2_MAC_IN, 2>(y1);
}
// ---- phase 3a: rows of one slice into a scratch tile ----
TFHE_HD void t2_inv_store(int lane, const uint32_t (&y)[32], uint32_t* T) {
#pragma unroll
    for (int q = 0; q < 8; q++)
        *reinterpret_cast<uint4*>(T + xs_chunk(lane, q)) = make_uint4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
}
// ---- phase 3b: lane = column.  Column pass of the inverse transform, exact signed lift; z[r] += value << sh ----
// The lift is exact when |true slice value| < p/2 (9.8 sigma for honestly generated keys); `nearest` tracks the
// smallest distance of a reduced value from the wrap point (p-1)/2 so that a caller can tell how much margin was left.
TFHE_HD void t2_inv_cols(int lane, const uint32_t* T, int sh, uint32_t (&z)[32]) {
    uint32_t x[32];
#pragma unroll
    for (int r = 0; r < 32; r++) x[r] = T[xs(r, lane)];
    gs32_lazy(x, TwUniform<true>());
    gs_norm<1>(x);   // [0,p)
#pragma unroll
    for (int r = 0; r < 32; r++) z[r] += (uint32_t)lift(x[r]) << sh;
}

// ---- key transform into the throughput layout: row pass + fold 2^32/N, chunk q of (row j, slice s) at dst + q * T2_Q_STRIDE ----
TFHE_HD void key_rows_t2(int lane, uint32_t* S, const uint32_t* twF, uint32_t* dst /* t2_bk_off(i, pw, 0, j, s, 0) */) {
    uint32_t x[32];
    fwd_rows(lane, S, twF, x);
#pragma unroll
    for (int c = 0; c < 32; c++) x[c] = csub(shoup_mul(x[c], NTT_MONT_NINV, NTT_MONT_NINV_SHOUP), P);
#pragma unroll
    for (int q = 0; q < 8; q++)
        *reinterpret_cast<uint4*>(dst + q * T2_Q_STRIDE + lane * 4) = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
}

}  // namespace tfhe
