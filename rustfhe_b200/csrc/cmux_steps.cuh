// cmux_steps.cuh -- the per-warp pieces of one CMUX step (TRGSW (x) TRLWE external product + accumulate).
//
// Reference semantics: hom_nand/src/tfhe.rs:103-110 (one fold step of blind_rotate), hom_nand/src/trgsw.rs:264-306
// (TRGSWRepF::cross), utils/src/math.rs:85-113 (rotate), utils/src/math.rs:300-326 (decomposition_i32_).
//
// Work split of one gate: 6 warps.  Warp w = 3*poly + k.
//   phase 1 (digit side)  : the three warps of a polynomial share the work of u = ((X^abar * acc - acc)[poly] + mask) ^ mask
//                           (a third of the rows each); then warp (poly,k) takes digit k of u, forward-NTTs it and leaves
//                           the spectrum in plane dh[3*poly+k]        (decomposition order = b-digits first, F10)
//   phase 2 (key side)    : warp (poly,k) computes sum_j dh[j] * BKpart_k[poly][j] (11-bit key slice k), inverse-NTTs
//                           it, lifts to the exact integer, shifts by 11k and leaves it in plane sp[3*poly+k]
//   phase 3               : acc[poly] += sp[3*poly+0] + sp[3*poly+1] + sp[3*poly+2]      (mod 2^32)
// Each function below is what ONE lane does between two warp-level synchronisation points; the CUDA kernel calls
// them with real shared memory and __syncwarp(), the CPU emulation (host_emul.cu) loops over the 32 lanes.
#pragma once
#include "ntt32.cuh"

namespace tfhe {

constexpr int NPOLY = 1024;
constexpr int LWE_N = 635;
constexpr int GADGET_L = 3;
constexpr int BK_ROWS = 6;                               // 2 * L
constexpr size_t BK_POLY_WORDS = 1024;                   // one transformed key-slice polynomial
constexpr size_t BK_SLAB_WORDS = BK_ROWS * 1024;         // [j][q][lane][4] : what one phase-2 warp streams per step
constexpr size_t BK_STEP_WORDS = 2 * 3 * BK_SLAB_WORDS;  // [poly][part] slabs of one TRGSW = 36 polys = 147,456 B (three key slices)
TFHE_HD size_t bk_step_words(int ns) { return (size_t)2 * ns * BK_SLAB_WORDS; }   // ns key slices

// device BK layout: word offset of (step i, output poly, key slice part, row j, chunk q, lane, e)
TFHE_HD size_t bk_off(int i, int poly, int part, int j, int q, int lane, int ns = 3) {
    return ((((((size_t)i * 2 + poly) * ns + part) * BK_ROWS + j) * 8 + q) * 32 + lane) * 4;
}

// value of (X^abar * A - A)[k], abar in [0, 2048).  Written with explicit sign masks (no predicated negate) so that every
// operation is a LOP3 / SHF / 3-input IADD3 on the ALU pipe: the FMA-heavy pipe is the one that bounds the kernel.
TFHE_HD uint32_t rot_diff(const uint32_t* A, uint32_t k, uint32_t abar) {
    const uint32_t ap = abar & 1023u;
    const uint32_t flip = 0u - (abar >> 10);                 // ~0 when abar >= N: one more factor X^N = -1
    const uint32_t t = k - ap;                               // wraps below zero when the rotated index folds around
    const uint32_t m = (uint32_t)((int32_t)t >> 31) ^ flip;  // ~0 -> take -A[...]:  -x = (x ^ m) - m
    const uint32_t rv = A[t & 1023u];
    return (rv ^ m) - m - A[k];
}

// ---- phase 1u: the masked source polynomial u = (src + mask) ^ mask is the same for the three digit warps of a
// polynomial, so each of them computes the rows r = part, part+3, ... of it and leaves them in the plain tile U ----
// ROTATE=true : src = X^abar * A - A   (blind rotation step);  ROTATE=false : src = A (plain external product)
template <bool ROTATE>
TFHE_HD void p1u(int lane, const uint32_t* A, uint32_t abar, uint32_t mask, int part, uint32_t* U) {
#pragma unroll
    for (int t = 0; t < 11; t++) {
        const int r = part + 3 * t;
        if (r < 32) {
            const uint32_t k = 32u * r + lane;
            const uint32_t src = ROTATE ? rot_diff(A, k, abar) : A[k];
            U[k] = add_alu(src, mask) ^ mask;
        }
    }
}
// digit `dw` (0 = most significant) of an already masked word, sign-extended from 6 bits
TFHE_HD int32_t masked_digit(uint32_t u, int dw) { return ((int32_t)(u << (6 * dw))) >> 26; }
// ---- phase 1a: lane = column c.  Digit `dw` of column c from U, column NTT, scatter into tile S ----
// Stage 0 pairs rows r and r+16 with the single twiddle psi^512; the digit of row r+16 goes through the table
// digit_tab[d + 32] = d * psi^512 mod p instead of a Shoup multiplication: a = X + T, b = X - T + p, both < 2p.
TFHE_HD void p1a(int lane, const uint32_t* U, int dw, uint32_t* S, const uint32_t* digit_tab) {
    uint32_t x[32];
#pragma unroll
    for (int r = 0; r < 16; r++) {
        const uint32_t X = to_residue(masked_digit(U[32 * r + lane], dw));
        const uint32_t T = digit_tab[masked_digit(U[32 * (r + 16) + lane], dw) + 32];
        x[r] = add_alu(X, T);
        x[r + 16] = X - T + P;
    }
    ct32_after_stage0(x, TwUniform<false>());
#pragma unroll
    for (int r = 0; r < 32; r++) S[swz(r, lane)] = x[r];
}
// generic forward column pass for an already prepared residue column (used by the key transform / poly-mul entry)
TFHE_HD void fwd_cols(int lane, uint32_t (&x)[32], uint32_t* S) {
    ct32(x, TwUniform<false>());
#pragma unroll
    for (int r = 0; r < 32; r++) S[swz(r, lane)] = x[r];
}
// ---- phase 1b: lane = row r.  Row NTT with this lane's twiddle row, full reduction, store back (row layout) ----
TFHE_HD void fwd_rows(int lane, uint32_t* S, const uint32_t* twF, uint32_t (&x)[32]) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const uint4 v = *reinterpret_cast<const uint4*>(S + swz_chunk(lane, q));
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
    ct32_wide(x, TwRow{twF + lane * TWB_STRIDE});   // column pass leaves < 8p: the row pass corrects inside stages 0, 2 and 4
#pragma unroll
    for (int c = 0; c < 32; c++) x[c] = csub(csub(csub(x[c], 2u * P2), P2), P);  // < 6p -> [0,p)
}
TFHE_HD void p1b(int lane, uint32_t* S, const uint32_t* twF) {
    uint32_t x[32];
    fwd_rows(lane, S, twF, x);
#pragma unroll
    for (int q = 0; q < 8; q++)
        *reinterpret_cast<uint4*>(S + swz_chunk(lane, q)) = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
}

// ---- phase 2a: lane = row r.  MAC over the 6 digit spectra against this warp's key slab, row INTT, store ----
// dh: 6 planes of 1024 words (row layout written by p1b); slab: BK_SLAB_WORDS words for (step, poly, part)
TFHE_HD void p2a_mac(int lane, const uint32_t* slab, const uint32_t* dh, uint32_t (&x)[32]) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
        uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int j = 0; j < BK_ROWS; j++) {
            const uint4 d = *reinterpret_cast<const uint4*>(dh + j * TILE_WORDS + swz_chunk(lane, q));
#if defined(__CUDA_ARCH__)
            const uint4 b = __ldg(reinterpret_cast<const uint4*>(slab) + (j * 8 + q) * 32 + lane);
#else
            const uint4 b = *(reinterpret_cast<const uint4*>(slab) + (j * 8 + q) * 32 + lane);
#endif
            a0 += (uint64_t)d.x * b.x; a1 += (uint64_t)d.y * b.y; a2 += (uint64_t)d.z * b.z; a3 += (uint64_t)d.w * b.w;
        }
        x[4 * q] = redc64(a0); x[4 * q + 1] = redc64(a1); x[4 * q + 2] = redc64(a2); x[4 * q + 3] = redc64(a3);
    }
}
// same, with the first two inverse row stages of each chunk folded into the MAC loop (see gs32_head4): butterfly work
// fills the wait for the next chunk's loads.  Finish with gs32_tail.
// dhb / dha: the three digit spectra of the b polynomial (key rows 0..2) and of the a polynomial (rows 3..5).
// SLAB_IN_SMEM: the slab was staged in shared memory (latency shape) instead of being streamed from global memory.
template <bool SLAB_IN_SMEM = false>
TFHE_HD void p2a_mac_head(int lane, const uint32_t* slab, const uint32_t* dhb, const uint32_t* dha, const uint32_t* twI, uint32_t (&x)[32]) {
    const TwRow tw{twI + lane * TWB_STRIDE};
#pragma unroll
    for (int q = 0; q < 8; q++) {
        uint64_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int j = 0; j < BK_ROWS; j++) {
            const uint4 d = *reinterpret_cast<const uint4*>((j < 3 ? dhb + j * TILE_WORDS : dha + (j - 3) * TILE_WORDS) + swz_chunk(lane, q));
#if defined(__CUDA_ARCH__)
            const uint4 b = SLAB_IN_SMEM ? *(reinterpret_cast<const uint4*>(slab) + (j * 8 + q) * 32 + lane)
                                         : __ldg(reinterpret_cast<const uint4*>(slab) + (j * 8 + q) * 32 + lane);
#else
            const uint4 b = *(reinterpret_cast<const uint4*>(slab) + (j * 8 + q) * 32 + lane);
#endif
            a0 += (uint64_t)d.x * b.x; a1 += (uint64_t)d.y * b.y; a2 += (uint64_t)d.z * b.z; a3 += (uint64_t)d.w * b.w;
        }
        x[4 * q] = redc64(a0); x[4 * q + 1] = redc64(a1); x[4 * q + 2] = redc64(a2); x[4 * q + 3] = redc64(a3);
        gs32_head4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3], q, tw);
    }
}
// latency shape, split MAC: the three key rows that meet this CTA's OWN digit spectra are accumulated first (64-bit partial
// sums for all 32 positions stay in registers) while the cluster barrier that delivers the peer's spectra is still pending;
// a second call adds the other three rows; p2a_mac_redc reduces.  j0: first key row of the part.
TFHE_HD void p2a_mac_part(int lane, const uint32_t* slab, const uint32_t* dh3, int j0, uint64_t (&acc)[32], bool first) {
#pragma unroll
    for (int q = 0; q < 8; q++) {
        uint64_t a0 = first ? 0 : acc[4 * q], a1 = first ? 0 : acc[4 * q + 1], a2 = first ? 0 : acc[4 * q + 2], a3 = first ? 0 : acc[4 * q + 3];
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const uint4 d = *reinterpret_cast<const uint4*>(dh3 + j * TILE_WORDS + swz_chunk(lane, q));
            const uint4 b = *(reinterpret_cast<const uint4*>(slab) + ((j0 + j) * 8 + q) * 32 + lane);
            a0 += (uint64_t)d.x * b.x; a1 += (uint64_t)d.y * b.y; a2 += (uint64_t)d.z * b.z; a3 += (uint64_t)d.w * b.w;
        }
        acc[4 * q] = a0; acc[4 * q + 1] = a1; acc[4 * q + 2] = a2; acc[4 * q + 3] = a3;
    }
}
TFHE_HD void p2a(int lane, const uint32_t* slab, const uint32_t* dh, const uint32_t* twI, uint32_t* S) {
    uint32_t x[32];
    p2a_mac_head(lane, slab, dh, dh + 3 * TILE_WORDS, twI, x);
    gs32_tail(x, TwRow{twI + lane * TWB_STRIDE});
    gs_norm<2>(x);
#pragma unroll
    for (int q = 0; q < 8; q++)
        *reinterpret_cast<uint4*>(S + swz_chunk(lane, q)) = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
}
// inverse row pass for a row already in registers (poly-mul entry)
TFHE_HD void inv_rows(int lane, uint32_t (&x)[32], const uint32_t* twI, uint32_t* S) {
    gs32(x, TwRow{twI + lane * TWB_STRIDE});
#pragma unroll
    for (int q = 0; q < 8; q++)
        *reinterpret_cast<uint4*>(S + swz_chunk(lane, q)) = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
}
// ---- phase 2b: lane = column c.  Column INTT -> exact signed integers, shifted by 11*part (kept in registers) ----
TFHE_HD void p2b(int lane, const uint32_t* S, int part, uint32_t (&x)[32], int ns = 3) {
#pragma unroll
    for (int r = 0; r < 32; r++) x[r] = S[swz(r, lane)];
    gs32_lazy(x, TwUniform<true>());
    gs_norm<1>(x);   // [0,p)
    const int sh = slice_shift(ns) * part;
#pragma unroll
    for (int r = 0; r < 32; r++) x[r] = (uint32_t)lift(x[r]) << sh;
}
// ---- latency shape: both passes of a transform through ONE code body ----
// The cluster kernel runs one warp per sub-partition, so its unrolled step is fetched once per step and warp: at 47 KB it
// does not fit the SM's 32 KB instruction cache and every step streams its code from L2 (7-13 % of the cycles by ncu, and the
// slower boxes of the pool pay more).  Here the column pass and the row pass of a transform share one 32-point network (a
// loop of two trips that must not be unrolled): the twiddles of either pass come from shared memory (tw_col: the
// warp-uniform column twiddles, 64 words; tw_rows: the per-lane rows), the column pass gives up the digit table and the
// constant-bank operands (about 2 % more work) and the step shrinks by a third.
// Forward: digit `dw` of column `lane` of U -> spectrum row `lane` in [0, p), in x; S is the transpose scratch.
TFHE_HD void fwd_shared(int lane, const uint32_t* U, int dw, uint32_t* S, const uint32_t* tw_col, const uint32_t* tw_rows, uint32_t (&x)[32]) {
#pragma unroll
    for (int r = 0; r < 32; r++) x[r] = to_residue(masked_digit(U[32 * r + lane], dw));
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        ct32_wide(x, TwRow{pass ? tw_rows + lane * TWB_STRIDE : tw_col});   // inputs < 8p (digits < p, column outputs < 6p) -> < 6p
        if (pass == 0) {
#pragma unroll
            for (int r = 0; r < 32; r++) S[swz(r, lane)] = x[r];
#if defined(__CUDA_ARCH__)
            __syncwarp();
#endif
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint4 v = *reinterpret_cast<const uint4*>(S + swz_chunk(lane, q));
                x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
            }
#if defined(__CUDA_ARCH__)
            __syncwarp();   // every lane has its row before the caller overwrites S with the spectrum
#endif
        }
    }
#pragma unroll
    for (int c = 0; c < 32; c++) x[c] = csub(csub(csub(x[c], 2u * P2), P2), P);  // < 6p -> [0,p)
}
// Inverse: x = row `lane` of the pointwise products, values < 2p -> exact signed slice values of column `lane`, shifted by
// the slice position (what p2a's row pass + p2b produce); T is the transpose scratch.
TFHE_HD void inv_shared(int lane, uint32_t (&x)[32], uint32_t* T, const uint32_t* tw_rows, const uint32_t* tw_col, int part, int ns) {
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {
        gs32_lazy(x, TwRow{pass ? tw_col : tw_rows + lane * TWB_STRIDE});
        if (pass == 0) {
            gs_norm<2>(x);
#pragma unroll
            for (int q = 0; q < 8; q++)
                *reinterpret_cast<uint4*>(T + swz_chunk(lane, q)) = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
#if defined(__CUDA_ARCH__)
            __syncwarp();
#endif
#pragma unroll
            for (int r = 0; r < 32; r++) x[r] = T[swz(r, lane)];
        }
    }
    gs_norm<1>(x);   // [0,p)
    const int sh = slice_shift(ns) * part;
#pragma unroll
    for (int r = 0; r < 32; r++) x[r] = (uint32_t)lift(x[r]) << sh;
}
// reduction of the 64-bit pointwise sums of the split MAC (p2a_mac_part x 2) without the fused inverse stages
TFHE_HD void p2a_mac_redc(const uint64_t (&acc)[32], uint32_t (&x)[32]) {
#pragma unroll
    for (int c = 0; c < 32; c++) x[c] = redc64(acc[c]);
}
// ---- phase 2c: plain (unswizzled) store: coefficient k = 32 r + lane ----
TFHE_HD void p2c(int lane, uint32_t* S, const uint32_t (&x)[32]) {
#pragma unroll
    for (int r = 0; r < 32; r++) S[32 * r + lane] = x[r];
}

// ---- key transform: column `lane` of key slice `part` of a torus polynomial ----
TFHE_HD void key_cols(int lane, const uint32_t* src, int part, uint32_t* S, int ns = 3) {
    uint32_t x[32];
#pragma unroll
    for (int r = 0; r < 32; r++) x[r] = to_residue(key_slice(src[32 * r + lane], part, ns));
    fwd_cols(lane, x, S);
}
// row pass + fold 2^32/N (Montgomery factor and the inverse transform's 1/N), store in the device BK layout
TFHE_HD void key_rows(int lane, uint32_t* S, const uint32_t* twF, uint32_t* dst /* one BK poly: [q][lane][4] */) {
    uint32_t x[32];
    fwd_rows(lane, S, twF, x);
#pragma unroll
    for (int c = 0; c < 32; c++) x[c] = csub(shoup_mul(x[c], NTT_MONT_NINV, NTT_MONT_NINV_SHOUP), P);
#pragma unroll
    for (int q = 0; q < 8; q++)
        *reinterpret_cast<uint4*>(dst + (q * 32 + lane) * 4) = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
}

}  // namespace tfhe
