/*
 * tfhe_oracle.c -- CPU oracle for the bootstrapped-HomNAND path (TEST INFRASTRUCTURE, see tfhe_oracle.h).
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 * Nothing here is copied from the reference: the semantics were re-derived.
 */
#define _GNU_SOURCE
#include "tfhe_oracle.h"
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define n_ ORC_n
#define N_ ORC_N
#define L_ ORC_L
#define KT_ ORC_KS_T

/* ------------------------------------------------------------------------------------------------
 * Seeded counter-based generator.  The reference draws everything from rand::thread_rng() and has no
 * seeding API (utils/src/math.rs:421,429,449,457,476) so the harness owns the generator.
 * ------------------------------------------------------------------------------------------------ */
static inline uint64_t fmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
uint64_t orc_rnd64(uint64_t seed, uint64_t stream, uint64_t idx) {
    uint64_t h = fmix64(seed ^ fmix64(stream * 0xD6E8FEB86659FD93ull + 0x1234567ull));
    return fmix64(h + idx * 0x9E3779B97F4A7C15ull);
}
/* round(N(0,alpha) * 2^32) as a wrapping torus increment (reference: Normal<f32> -> torus!, math.rs:411-432).
 * Box-Muller evaluated with correctly rounded IEEE-754 double operations only (+ - * / sqrt, fixed order, no libm, no
 * fused multiply-add: this file is compiled as ISO C, where GCC does not contract, and the volatile temporaries pin
 * it), so that any conforming CPU -- and the product's device keygen -- produce the same bits from the same seed. */
static inline double o_mul(double a, double b) { volatile double r = a * b; return r; }
static inline double o_add(double a, double b) { volatile double r = a + b; return r; }
static inline double o_div(double a, double b) { volatile double r = a / b; return r; }
static double o_log(double u) { /* u normal, in (0,1] */
    uint64_t b; memcpy(&b, &u, 8);
    int e = (int)((b >> 52) & 0x7FF) - 1022;
    uint64_t mb = (b & 0x000FFFFFFFFFFFFFull) | 0x3FE0000000000000ull;
    double m; memcpy(&m, &mb, 8);
    if (m < 0.70710678118654752440) { m = o_mul(m, 2.0); e -= 1; }
    const double s = o_div(o_add(m, -1.0), o_add(m, 1.0));
    const double z = o_mul(s, s);
    double p = 1.0 / 23.0;
    for (int k = 21; k >= 1; k -= 2) p = o_add(o_mul(p, z), 1.0 / (double)k);
    return o_add(o_mul((double)e, 0.69314718055994530942), o_mul(o_mul(2.0, s), p));
}
static double o_cos2pi(double v) { /* v in [0,1) */
    static const double CF[10] = {-1.0 / 6402373705728000.0, 1.0 / 20922789888000.0, -1.0 / 87178291200.0, 1.0 / 479001600.0,
                                  -1.0 / 3628800.0, 1.0 / 40320.0, -1.0 / 720.0, 1.0 / 24.0, -0.5, 1.0};
    static const double SF[10] = {-1.0 / 121645100408832000.0, 1.0 / 355687428096000.0, -1.0 / 1307674368000.0, 1.0 / 6227020800.0,
                                  -1.0 / 39916800.0, 1.0 / 362880.0, -1.0 / 5040.0, 1.0 / 120.0, -1.0 / 6.0, 1.0};
    const double t = o_mul(v, 4.0);
    const int q = (int)t;
    double r = o_add(t, -(double)q);
    const int fold = r > 0.5;
    if (fold) r = o_add(1.0, -r);
    const double x = o_mul(r, 1.57079632679489661923), z = o_mul(x, x);
    double c = CF[0], s = SF[0];
    for (int k = 1; k < 10; k++) { c = o_add(o_mul(c, z), CF[k]); s = o_add(o_mul(s, z), SF[k]); }
    s = o_mul(s, x);
    const double cq = fold ? s : c, sq = fold ? c : s;
    return q == 0 ? cq : q == 1 ? -sq : q == 2 ? -cq : sq;
}
int32_t orc_gauss_torus(uint64_t seed, uint64_t stream, uint64_t idx, double alpha) {
    double u1 = o_mul((double)((orc_rnd64(seed, stream, 2 * idx) >> 11) + 1), 1.0 / 9007199254740992.0);
    double u2 = o_mul((double)(orc_rnd64(seed, stream, 2 * idx + 1) >> 11), 1.0 / 9007199254740992.0);
    double g = o_mul(sqrt(o_mul(-2.0, o_log(u1))), o_cos2pi(u2));
    return (int32_t)llrint(o_mul(g, alpha * 4294967296.0));
}
enum { ST_S0 = 1, ST_S1 = 2, ST_BK_A = 3, ST_BK_E = 4, ST_KSK_A = 5, ST_KSK_E = 6, ST_ENC_A = 7, ST_ENC_E = 8 };

void orc_keygen_secret(uint64_t seed, uint8_t* s0, uint8_t* s1) {
    for (int i = 0; i < n_; i++) s0[i] = (uint8_t)(orc_rnd64(seed, ST_S0, i) >> 63);
    for (int i = 0; i < N_; i++) s1[i] = (uint8_t)(orc_rnd64(seed, ST_S1, i) >> 63);
}

/* ------------------------------------------------------------------------------------------------
 * Torus32 helpers
 * ------------------------------------------------------------------------------------------------ */
/* utils/src/math.rs:691-696 : Decimal(((val - val.floor()).fract() * (u32::MAX as f32)) as u32); u32::MAX as f32 == 2^32;
 * Rust `as u32` saturates. */
uint32_t orc_torus_from_f32(float v) {
    float f = v - floorf(v);
    f = f - truncf(f);
    float x = f * 4294967296.0f;
    if (!(x > 0.0f)) return 0u;
    if (x >= 4294967296.0f) return 0xFFFFFFFFu;
    return (uint32_t)x;
}
/* utils/src/math.rs:684-690 : (self.0 as f32) * (1.0 / (u32::MAX as f32)) */
float orc_torus_to_f32(uint32_t t) { return (float)t * (1.0f / 4294967296.0f); }

/* ------------------------------------------------------------------------------------------------
 * Polynomial glue
 * ------------------------------------------------------------------------------------------------ */
/* Polynomial::rotate, utils/src/math.rs:85-113 : multiply by X^k mod (X^n + 1), k taken mod 2n (floor mod). */
void orc_rotate(const uint32_t* p, int n, int k, uint32_t* out) {
    int m = k % (2 * n);
    if (m < 0) m += 2 * n;
    int flip = 0;
    if (m > n) { m -= n; flip = 1; } /* for m > n the reference negates the (m-n) rotation */
    for (int i = 0; i < n; i++) {
        uint32_t v = (i < m) ? (0u - p[n - m + i]) : p[i - m];
        out[i] = flip ? (0u - v) : v;
    }
}
/* Cross for Polynomial (schoolbook negacyclic), utils/src/math.rs:238-255 + convolution 701-723 */
void orc_negacyclic_mul_schoolbook(const uint32_t* a, const int32_t* d, int n, uint32_t* out) {
    for (int k = 0; k < n; k++) {
        uint32_t acc = 0;
        for (int i = 0; i <= k; i++) acc += a[i] * (uint32_t)d[k - i];
        for (int i = k + 1; i < n; i++) acc -= a[i] * (uint32_t)d[n + k - i];
        out[k] = acc;
    }
}
/* Torus32::make_decomp_mask, utils/src/math.rs:542-560 -- note the rounding bit is ADDED twice when total != l*bits
 * (line 546 and the i == l iteration of the loop at 548-551): (3,6) -> 0x02084000. */
uint32_t orc_make_decomp_mask(uint32_t l, uint32_t bits) {
    const uint32_t total = 32;
    uint32_t u = 0;
    if (total - l * bits != 0) {
        u += 1u << (total - l * bits - 1);
        for (uint32_t i = l; i >= 1; i--) u += 1u << (total - i * bits - 1);
    } else {
        for (uint32_t i = l - 1; i >= 1; i--) u += 1u << (total - i * bits - 1);
    }
    return u;
}
/* the OR-built mask inlined in Torus32::decomposition_i32, utils/src/math.rs:582-591 -- the one the KATs pin */
uint32_t orc_tested_decomp_mask(uint32_t l, uint32_t bits) {
    const uint32_t total = 32;
    uint32_t u = 0;
    uint32_t top = (total - l * bits != 0) ? l : l - 1;
    for (uint32_t i = 1; i <= top; i++) u |= 1u << (total - i * bits - 1);
    return u;
}
/* Torus32::decomposition_i32_, utils/src/math.rs:561-577 */
void orc_decompose_scalar(uint32_t x, uint32_t l, uint32_t bits, uint32_t mask, int32_t* out) {
    uint32_t u = (x + mask) ^ mask;
    uint32_t m = (1u << bits) - 1;
    for (uint32_t i = 0; i < l; i++) {
        uint32_t v = (u >> (32 - bits * (i + 1))) & m;
        out[i] = (int32_t)((v & (1u << (bits - 1))) * 0xfffffffeu + v);
    }
}
/* Torus32::decomposition_u32, utils/src/math.rs:596-614 : unsigned digits after adding the rounding bit */
void orc_decompose_u32_scalar(uint32_t x, uint32_t l, uint32_t bits, uint32_t* out) {
    uint32_t u = x + ((32 - l * bits != 0) ? (1u << (32 - l * bits - 1)) : 0u);
    uint32_t m = (1u << bits) - 1;
    for (uint32_t i = 0; i < l; i++) out[i] = (u >> (32 - bits * (i + 1))) & m;
}
/* Polynomial::decomposition_i32_, utils/src/math.rs:300-326 : out[i][k] = digit i of coefficient k */
void orc_decompose(const uint32_t* p, uint32_t mask, int32_t* out) {
    int32_t d[L_];
    for (int k = 0; k < N_; k++) {
        orc_decompose_scalar(p[k], L_, ORC_BGBIT, mask, d);
        for (int i = 0; i < L_; i++) out[i * N_ + k] = d[i];
    }
}

/* ------------------------------------------------------------------------------------------------
 * TLWE level-0: encode / encrypt / phase / decrypt  (hom_nand/src/tlwe.rs:181-240)
 * ------------------------------------------------------------------------------------------------ */
static inline uint32_t encode_bit(uint8_t b) { return b ? ORC_MU : (0u - ORC_MU); } /* tlwe.rs:181-186 */
/* tlwe.rs:187-194 : f32(phase) < 0.5 -> One */
static inline uint8_t decode_phase(uint32_t ph) { return orc_torus_to_f32(ph) < 0.5f ? 1 : 0; }

/* tlwe.rs:213-228 : b = sum_{s_i=1} a_i + e + m.  Masks are full 32-bit uniform (deliberately not the
 * reference's 24-bit Uniform<f32> quirk); noise alpha = 2^-15 (tlwe.rs:176). */
void orc_tlwe_encrypt_bits(uint64_t seed, uint64_t ct0, const uint8_t* s0, const uint8_t* bits, size_t B, uint32_t* out) {
    for (size_t g = 0; g < B; g++) {
        uint32_t* ct = out + g * (n_ + 1);
        uint64_t id = ct0 + g;
        uint32_t b = encode_bit(bits[g]) + (uint32_t)orc_gauss_torus(seed, ST_ENC_E, id, 1.0 / 32768.0);
        for (int i = 0; i < n_; i++) {
            uint32_t a = (uint32_t)(orc_rnd64(seed, ST_ENC_A, id * n_ + i) >> 32);
            ct[1 + i] = a;
            if (s0[i]) b += a;
        }
        ct[0] = b;
    }
}
/* tlwe.rs:230-240 */
void orc_tlwe_phase(const uint8_t* s0, const uint32_t* ct, size_t B, uint32_t* phase) {
    for (size_t g = 0; g < B; g++) {
        const uint32_t* c = ct + g * (n_ + 1);
        uint32_t acc = c[0];
        for (int i = 0; i < n_; i++)
            if (s0[i]) acc -= c[1 + i];
        phase[g] = acc;
    }
}
void orc_tlwe_decrypt_bits(const uint8_t* s0, const uint32_t* ct, size_t B, uint8_t* bits) {
    for (size_t g = 0; g < B; g++) {
        uint32_t ph;
        orc_tlwe_phase(s0, ct + g * (n_ + 1), 1, &ph);
        bits[g] = decode_phase(ph);
    }
}
void orc_tlwe1_phase(const uint8_t* s1, const uint32_t* ct, size_t B, uint32_t* phase) {
    for (size_t g = 0; g < B; g++) {
        const uint32_t* c = ct + g * (N_ + 1);
        uint32_t acc = c[0];
        for (int i = 0; i < N_; i++)
            if (s1[i]) acc -= c[1 + i];
        phase[g] = acc;
    }
}

/* gate pre-combinations, hom_nand/src/tfhe.rs:27-71 (+ TLWERep Add/Sub/Neg/Mul, tlwe.rs:88-159) */
void orc_gate_linear(int op, const uint32_t* in0, const uint32_t* in1, size_t B, uint32_t* out) {
    for (size_t g = 0; g < B; g++) {
        const uint32_t* x = in0 + g * (n_ + 1);
        const uint32_t* y = in1 ? in1 + g * (n_ + 1) : x;
        uint32_t* o = out + g * (n_ + 1);
        for (int i = 0; i <= n_; i++) {
            uint32_t v;
            switch (op) {
            case ORC_NAND: v = 0u - (x[i] + y[i]); break;                /* trivial(1/8) - (c0+c1)     tfhe.rs:41-47 */
            case ORC_AND: v = x[i] + y[i]; break;                        /* (c0+c1) - trivial(1/8)     tfhe.rs:48-54 */
            case ORC_OR: v = x[i] + y[i]; break;                         /* (c0+c1) + trivial(1/8)     tfhe.rs:55-61 */
            case ORC_XOR: v = 2u * (x[i] + y[i]); break;                 /* (c0+c1)*2 + trivial(1/4)   tfhe.rs:62-68 */
            case ORC_NOT: v = 0u - x[i]; break;                          /* -c                         tfhe.rs:69-71 */
            case ORC_ANDNY: v = y[i] - x[i]; break;                      /* hom_and(-control, input_0) tfhe.rs:34    */
            default: v = x[i]; break;
            }
            o[i] = v;
        }
        switch (op) {
        case ORC_NAND: o[0] += ORC_MU; break;
        case ORC_AND: o[0] -= ORC_MU; break;
        case ORC_OR: o[0] += ORC_MU; break;
        case ORC_XOR: o[0] += 2u * ORC_MU; break;
        case ORC_ANDNY: o[0] -= ORC_MU; break;
        default: break;
        }
    }
}

/* TRLWERep::sample_extract_index, hom_nand/src/trlwe.rs:110-121 */
void orc_sample_extract(const uint32_t* trlwe, int index, uint32_t* out) {
    const uint32_t* b = trlwe;
    const uint32_t* a = trlwe + N_;
    out[0] = b[index];
    for (int i = 0; i < N_; i++) out[1 + i] = (i <= index) ? a[index - i] : (0u - a[N_ + index - i]);
}
void orc_sample_extract0(const uint32_t* trlwe, uint32_t* out) { orc_sample_extract(trlwe, 0, out); }

/* digits of identity_key_switch, hom_nand/src/tlwe.rs:47-64 : u = a_i + 2^15 ; digit_l = (u >> (32 - 2(l+1))) & 3.
 * The 8 two-bit digits are exactly the top 16 bits of u. */
void orc_ks_digits(const uint32_t* lwe1, uint16_t* dig) {
    for (int i = 0; i < N_; i++) dig[i] = (uint16_t)((lwe1[1 + i] + 0x8000u) >> 16);
}
/* TLWERep::identity_key_switch, hom_nand/src/tlwe.rs:43-73.  ksk layout [N][T][3][n+1] (digit value t=1..3 at index t-1;
 * the reference stores an unreachable 4th entry, tlwe.rs:243-245 / SURVEY F8). */
void orc_key_switch(const uint32_t* ksk, const uint32_t* lwe1, uint32_t* out) {
    memset(out, 0, (n_ + 1) * sizeof(uint32_t));
    out[0] = lwe1[0];
    for (int i = 0; i < N_; i++) {
        uint32_t u = lwe1[1 + i] + (1u << 15);
        for (int l = 0; l < KT_; l++) {
            uint32_t d = (u >> (32 - 2 * (l + 1))) & 3u;
            if (d != 0) {
                const uint32_t* row = ksk + (((size_t)i * KT_ + l) * 3 + (d - 1)) * (n_ + 1);
                for (int c = 0; c <= n_; c++) out[c] -= row[c];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Exact layer: an independent 2-prime CPU NTT (textbook twist + cyclic radix-2), CRT-lifted to the exact
 * signed integer, then reduced mod 2^32.  Result range of one external product: |x| < 6*1024*2^31*2^5 < 2^49
 * (SURVEY F5), product of the primes ~ 2^61.7.
 * ------------------------------------------------------------------------------------------------ */
#define Q1 2013265921ull /* 15*2^27+1 */
#define Q2 1811939329ull /* 27*2^26+1 */
typedef struct { uint32_t tw[N_], itw[N_], psi[N_], ipsi[N_]; uint32_t q; } ntt_tab;
static ntt_tab T1, T2;
static int tabs_ready = 0;

static uint64_t powmod(uint64_t b, uint64_t e, uint64_t q) {
    uint64_t r = 1; b %= q;
    while (e) { if (e & 1) r = (unsigned __int128)r * b % q; b = (unsigned __int128)b * b % q; e >>= 1; }
    return r;
}
static uint64_t find_psi(uint64_t q) { /* primitive 2N-th root of unity */
    for (uint64_t g = 2;; g++) {
        uint64_t c = powmod(g, (q - 1) / (2 * N_), q);
        if (powmod(c, N_, q) == q - 1) return c;
    }
}
static void build_tab(ntt_tab* t, uint64_t q) {
    t->q = (uint32_t)q;
    uint64_t psi = find_psi(q), ipsi = powmod(psi, q - 2, q);
    uint64_t w = psi * psi % q, iw = powmod(w, q - 2, q), ninv = powmod(N_, q - 2, q);
    uint64_t a = 1, b = 1, c = 1, d = ninv;
    for (int i = 0; i < N_; i++) {
        t->tw[i] = (uint32_t)a; t->itw[i] = (uint32_t)b; t->psi[i] = (uint32_t)c; t->ipsi[i] = (uint32_t)d;
        a = a * w % q; b = b * iw % q; c = c * psi % q; d = d * ipsi % q;
    }
}
static void ensure_tabs(void) {
#pragma omp critical(orc_tabs)
    if (!tabs_ready) { build_tab(&T1, Q1); build_tab(&T2, Q2); tabs_ready = 1; }
}
static inline uint32_t bitrev10(uint32_t x) { uint32_t r = 0; for (int i = 0; i < 10; i++) r |= ((x >> i) & 1u) << (9 - i); return r; }
/* cyclic NTT, in place, DIT with bit-reversal permutation first */
#define DEFINE_NTT(NAME, Q)                                                                    \
    static void NAME(uint32_t* a, const uint32_t* tw) {                                        \
        for (uint32_t i = 0; i < N_; i++) { uint32_t j = bitrev10(i); if (i < j) { uint32_t t = a[i]; a[i] = a[j]; a[j] = t; } } \
        for (int len = 2; len <= N_; len <<= 1) {                                              \
            int half = len >> 1, step = N_ / len;                                              \
            for (int s = 0; s < N_; s += len)                                                  \
                for (int k = 0; k < half; k++) {                                               \
                    uint64_t u = a[s + k], v = (uint64_t)a[s + k + half] * tw[k * step] % Q;   \
                    uint64_t x = u + v; if (x >= Q) x -= Q;                                    \
                    uint64_t y = u + Q - v; if (y >= Q) y -= Q;                                \
                    a[s + k] = (uint32_t)x; a[s + k + half] = (uint32_t)y;                     \
                }                                                                              \
        }                                                                                      \
    }
DEFINE_NTT(ntt_q1, Q1)
DEFINE_NTT(ntt_q2, Q2)
static inline uint32_t to_res(int64_t v, uint64_t q) { int64_t r = v % (int64_t)q; if (r < 0) r += q; return (uint32_t)r; }

/* forward negacyclic transform of signed integer poly into residue spectra */
static void fwd_signed(const int64_t* x, uint32_t* o1, uint32_t* o2) {
    for (int i = 0; i < N_; i++) {
        o1[i] = (uint32_t)((uint64_t)to_res(x[i], Q1) * T1.psi[i] % Q1);
        o2[i] = (uint32_t)((uint64_t)to_res(x[i], Q2) * T2.psi[i] % Q2);
    }
    ntt_q1(o1, T1.tw); ntt_q2(o2, T2.tw);
}
/* inverse + CRT lift to the centred integer, reduced mod 2^32 */
static void inv_to_torus(uint32_t* s1, uint32_t* s2, uint32_t* out) {
    ntt_q1(s1, T1.itw); ntt_q2(s2, T2.itw);
    static uint64_t q1inv_q2 = 0;
    if (!q1inv_q2) q1inv_q2 = powmod(Q1 % Q2, Q2 - 2, Q2);
    const unsigned __int128 M = (unsigned __int128)Q1 * Q2;
    for (int i = 0; i < N_; i++) {
        uint64_t r1 = (uint64_t)s1[i] * T1.ipsi[i] % Q1; /* ipsi already carries 1/N */
        uint64_t r2 = (uint64_t)s2[i] * T2.ipsi[i] % Q2;
        uint64_t t = (r2 + Q2 - r1 % Q2) % Q2 * q1inv_q2 % Q2;
        unsigned __int128 x = (unsigned __int128)r1 + (unsigned __int128)Q1 * t; /* in [0, M) */
        if (x > M / 2) out[i] = (uint32_t)(uint64_t)x - (uint32_t)(uint64_t)M; else out[i] = (uint32_t)(uint64_t)x;
    }
}
void orc_negacyclic_mul_ntt(const uint32_t* a, const int32_t* d, uint32_t* out) {
    ensure_tabs();
    int64_t xa[N_], xd[N_];
    uint32_t a1[N_], a2[N_], d1[N_], d2[N_];
    for (int i = 0; i < N_; i++) { xa[i] = (int32_t)a[i]; xd[i] = d[i]; } /* torus read as signed: fft_processor_spqlios.cpp:100-106 */
    fwd_signed(xa, a1, a2); fwd_signed(xd, d1, d2);
    for (int i = 0; i < N_; i++) { a1[i] = (uint32_t)((uint64_t)a1[i] * d1[i] % Q1); a2[i] = (uint32_t)((uint64_t)a2[i] * d2[i] % Q2); }
    inv_to_torus(a1, a2, out);
}

/* TRGSW (x) TRLWE external product, exact: hom_nand/src/trgsw.rs:264-306.
 * trgsw layout [row 0..2L)[poly 0=cipher(b),1=p_key(a)][N]; rows 0..L multiply the b-digits, rows L..2L the a-digits
 * (trgsw.rs:290-299). */
typedef struct { uint32_t* s1; uint32_t* s2; } bk_spec; /* [n][2L][2][N] per prime */
static void ext_prod_spec(const uint32_t* g1, const uint32_t* g2, const uint32_t* trlwe, uint32_t mask, uint32_t* out) {
    int32_t dig[2 * L_ * N_];
    orc_decompose(trlwe, mask, dig);                /* b digits -> rows 0..L  */
    orc_decompose(trlwe + N_, mask, dig + L_ * N_); /* a digits -> rows L..2L */
    uint32_t accb1[N_] = {0}, accb2[N_] = {0}, acca1[N_] = {0}, acca2[N_] = {0};
    int64_t x[N_];
    uint32_t d1[N_], d2[N_];
    for (int j = 0; j < 2 * L_; j++) {
        for (int k = 0; k < N_; k++) x[k] = dig[j * N_ + k];
        fwd_signed(x, d1, d2);
        const uint32_t* b1 = g1 + (size_t)(j * 2 + 0) * N_; const uint32_t* a1 = g1 + (size_t)(j * 2 + 1) * N_;
        const uint32_t* b2 = g2 + (size_t)(j * 2 + 0) * N_; const uint32_t* a2 = g2 + (size_t)(j * 2 + 1) * N_;
        for (int k = 0; k < N_; k++) {
            accb1[k] = (uint32_t)((accb1[k] + (uint64_t)d1[k] * b1[k]) % Q1);
            acca1[k] = (uint32_t)((acca1[k] + (uint64_t)d1[k] * a1[k]) % Q1);
            accb2[k] = (uint32_t)((accb2[k] + (uint64_t)d2[k] * b2[k]) % Q2);
            acca2[k] = (uint32_t)((acca2[k] + (uint64_t)d2[k] * a2[k]) % Q2);
        }
    }
    inv_to_torus(accb1, accb2, out);
    inv_to_torus(acca1, acca2, out + N_);
}
static void trgsw_to_spec(const uint32_t* trgsw, uint32_t* g1, uint32_t* g2) {
    int64_t x[N_];
    for (int r = 0; r < 2 * L_ * 2; r++) {
        for (int k = 0; k < N_; k++) x[k] = (int32_t)trgsw[(size_t)r * N_ + k];
        fwd_signed(x, g1 + (size_t)r * N_, g2 + (size_t)r * N_);
    }
}
void orc_external_product_exact(const uint32_t* trgsw, const uint32_t* trlwe, uint32_t mask, uint32_t* out) {
    ensure_tabs();
    uint32_t* g1 = malloc(sizeof(uint32_t) * 2 * L_ * 2 * N_);
    uint32_t* g2 = malloc(sizeof(uint32_t) * 2 * L_ * 2 * N_);
    trgsw_to_spec(trgsw, g1, g2);
    ext_prod_spec(g1, g2, trlwe, mask, out);
    free(g1); free(g2);
}
void* orc_exact_bk_prepare(const uint32_t* bk) {
    ensure_tabs();
    bk_spec* h = malloc(sizeof(bk_spec));
    size_t per = (size_t)2 * L_ * 2 * N_;
    h->s1 = malloc(sizeof(uint32_t) * per * n_);
    h->s2 = malloc(sizeof(uint32_t) * per * n_);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_; i++) trgsw_to_spec(bk + per * i, h->s1 + per * i, h->s2 + per * i);
    return h;
}
void orc_exact_bk_free(void* p) { bk_spec* h = p; if (!h) return; free(h->s1); free(h->s2); free(h); }

/* TFHE::blind_rotate, hom_nand/src/tfhe.rs:89-113 with cmux = cross(rot - acc) + acc (trgsw.rs:319-321).
 * bbar = b >> 21 (floor); abar_i = (a_i + 2^20) >> 21 (round)  (SURVEY F9). testvec = all 1/8 on b, 0 on a (tfhe.rs:85). */
typedef void (*extprod_fn)(const void* ctx, int i, const uint32_t* trlwe, uint32_t mask, uint32_t* out);
static void blind_rotate_generic(extprod_fn ep, const void* ctx, const uint32_t* tlwe, uint32_t mask, int nsteps, uint32_t* acc) {
    uint32_t tv[2 * N_], rot[2 * N_], diff[2 * N_], prod[2 * N_];
    for (int k = 0; k < N_; k++) { tv[k] = ORC_MU; tv[N_ + k] = 0; }
    int bbar = (int)(tlwe[0] >> 21);
    orc_rotate(tv, N_, -bbar, acc);
    orc_rotate(tv + N_, N_, -bbar, acc + N_);
    for (int i = 0; i < nsteps; i++) {
        int abar = (int)((tlwe[1 + i] + (1u << 20)) >> 21);
        orc_rotate(acc, N_, abar, rot);
        orc_rotate(acc + N_, N_, abar, rot + N_);
        for (int k = 0; k < 2 * N_; k++) diff[k] = rot[k] - acc[k];
        ep(ctx, i, diff, mask, prod);
        for (int k = 0; k < 2 * N_; k++) acc[k] += prod[k];
    }
}
static void ep_exact(const void* ctx, int i, const uint32_t* trlwe, uint32_t mask, uint32_t* out) {
    const bk_spec* h = ctx;
    size_t per = (size_t)2 * L_ * 2 * N_;
    ext_prod_spec(h->s1 + per * i, h->s2 + per * i, trlwe, mask, out);
}
void orc_blind_rotate_exact(const void* bkh, const uint32_t* tlwe, uint32_t mask, int nsteps, uint32_t* out) {
    blind_rotate_generic(ep_exact, bkh, tlwe, mask, nsteps, out);
}
/* TFHE::bootstrap, hom_nand/src/tfhe.rs:73-88 : blind_rotate -> sample_extract_index(0) -> identity_key_switch */
void orc_bootstrap_exact(const void* bkh, const uint32_t* ksk, const uint32_t* in, size_t B, uint32_t mask, uint32_t* out,
                         uint32_t* out_lwe1) {
#pragma omp parallel for schedule(dynamic)
    for (long g = 0; g < (long)B; g++) {
        uint32_t acc[2 * N_], lwe1[N_ + 1];
        orc_blind_rotate_exact(bkh, in + (size_t)g * (n_ + 1), mask, n_, acc);
        orc_sample_extract0(acc, lwe1);
        if (out_lwe1) memcpy(out_lwe1 + (size_t)g * (N_ + 1), lwe1, sizeof lwe1);
        orc_key_switch(ksk, lwe1, out + (size_t)g * (n_ + 1));
    }
}
void orc_trlwe_phase(const uint8_t* s1, const uint32_t* trlwe, uint32_t* phase) {
    int32_t s[N_];
    uint32_t as[N_];
    for (int i = 0; i < N_; i++) s[i] = s1[i];
    orc_negacyclic_mul_schoolbook(trlwe + N_, s, N_, as);
    for (int i = 0; i < N_; i++) phase[i] = trlwe[i] - as[i];
}

/* ------------------------------------------------------------------------------------------------
 * Key generation (host side in the reference: tfhe.rs:21-25,119-126; tlwe.rs:247-277; trgsw.rs:118-138,213-229;
 * trlwe.rs:127-137).  The a (x) s1 product is computed EXACTLY here (the reference uses its FFT, fft_cross
 * math.rs:337-347); noise alpha_bk = 2^-25 (trlwe.rs:77), alpha_lv0 = 2^-15 (tlwe.rs:176).
 * ------------------------------------------------------------------------------------------------ */
void orc_keygen_bk(uint64_t seed, const uint8_t* s0, const uint8_t* s1, uint32_t* bk) {
    ensure_tabs();
    int32_t s[N_];
    for (int i = 0; i < N_; i++) s[i] = s1[i];
#pragma omp parallel for schedule(static)
    for (int i = 0; i < n_; i++) {
        for (int j = 0; j < 2 * L_; j++) {
            uint32_t* B = bk + (((size_t)i * 2 * L_ + j) * 2 + 0) * N_;
            uint32_t* A = bk + (((size_t)i * 2 * L_ + j) * 2 + 1) * N_;
            uint64_t base = ((uint64_t)i * 2 * L_ + j) * N_;
            for (int k = 0; k < N_; k++) A[k] = (uint32_t)(orc_rnd64(seed, ST_BK_A, base + k) >> 32);
            orc_negacyclic_mul_ntt(A, s, B);
            for (int k = 0; k < N_; k++) B[k] += (uint32_t)orc_gauss_torus(seed, ST_BK_E, base + k, 1.0 / 33554432.0);
            /* TRGSW of the integer s0_i: rows 0..L get mu/Bg^(j+1) on cipher[0], rows L..2L on p_key[0] (trgsw.rs:213-229) */
            uint32_t mu = (uint32_t)s0[i] << (32 - ORC_BGBIT * ((j % L_) + 1));
            if (j < L_) B[0] += mu; else A[0] += mu;
        }
    }
}
void orc_keygen_ksk(uint64_t seed, const uint8_t* s0, const uint8_t* s1, uint32_t* ksk) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < N_; i++)
        for (int l = 0; l < KT_; l++)
            for (int t = 1; t <= 3; t++) {
                uint64_t rowid = ((uint64_t)i * KT_ + l) * 3 + (t - 1);
                uint32_t* row = ksk + rowid * (n_ + 1);
                /* message t * s1_i / 2^(2(l+1)) : tlwe.rs:253-258,269-273 */
                uint32_t b = ((uint32_t)(t * s1[i]) << (32 - ORC_KS_BASEBIT * (l + 1))) +
                             (uint32_t)orc_gauss_torus(seed, ST_KSK_E, rowid, 1.0 / 32768.0);
                for (int c = 0; c < n_; c++) {
                    uint32_t a = (uint32_t)(orc_rnd64(seed, ST_KSK_A, rowid * n_ + c) >> 32);
                    row[1 + c] = a;
                    if (s0[c]) b += a;
                }
                row[0] = b;
            }
}

/* ------------------------------------------------------------------------------------------------
 * Reference layer: the reference's own FFT library (oracle/_ref/libspqlios_ref.so) under restated glue.
 * ABI: utils/src/spqlios/spqlios-wrapper.cpp:4-53 (bound from Rust at utils/src/spqlios.rs:18-32).
 * ------------------------------------------------------------------------------------------------ */
typedef void* (*fn_new)(int32_t);
typedef void (*fn_ifft_u32)(void*, double*, const uint32_t*);
typedef void (*fn_ifft_i32)(void*, double*, const int32_t*);
typedef void (*fn_fft_u32)(void*, uint32_t*, const double*);
typedef void (*fn_poly_mul)(void*, uint32_t*, const uint32_t*, const uint32_t*);
static struct { void* dl; fn_new mk; fn_ifft_u32 ifft_u32; fn_ifft_i32 ifft_i32; fn_fft_u32 fft_u32; fn_poly_mul poly_mul; } R;
#define MAX_THREADS 256
static void* handles[MAX_THREADS]; /* one handle per thread: the handle owns scratch buffers (spqlios-fft.h:25-30) */

int orc_ref_init(const char* libpath) {
    if (R.dl) return 0;
    void* dl = dlopen(libpath, RTLD_NOW | RTLD_LOCAL);
    if (!dl) { fprintf(stderr, "orc_ref_init: %s\n", dlerror()); return 1; }
    R.mk = (fn_new)dlsym(dl, "Spqlios_new");
    R.ifft_u32 = (fn_ifft_u32)dlsym(dl, "Spqlios_ifft_u32");
    R.ifft_i32 = (fn_ifft_i32)dlsym(dl, "Spqlios_ifft_i32");
    R.fft_u32 = (fn_fft_u32)dlsym(dl, "Spqlios_fft_u32");
    R.poly_mul = (fn_poly_mul)dlsym(dl, "Spqlios_poly_mul");
    if (!R.mk || !R.ifft_u32 || !R.ifft_i32 || !R.fft_u32 || !R.poly_mul) return 2;
    R.dl = dl;
    return 0;
}
int orc_ref_available(void) { return R.dl != NULL; }
static void* ref_handle(void) {
    int t = 0;
#ifdef _OPENMP
    t = omp_get_thread_num();
#endif
    if (t >= MAX_THREADS) abort();
    /* only ever N=1024 in this process: `static const double _2sN` latches the first N (fft_processor_spqlios.cpp:110,158) */
    if (!handles[t]) {
#pragma omp critical(orc_ref_new)
        handles[t] = R.mk(N_);
    }
    return handles[t];
}
void orc_ref_poly_mul(const uint32_t* a, const uint32_t* b, uint32_t* out) { R.poly_mul(ref_handle(), out, a, b); }
void orc_ref_ifft_fft_roundtrip(const uint32_t* a, uint32_t* out) {
    double f[N_];
    void* h = ref_handle();
    R.ifft_u32(h, f, a);
    R.fft_u32(h, out, f);
}
/* spqlios.rs fft_test (utils/src/spqlios.rs:243-276) needs N=16; because of the latched 2/N call these only in
 * a process that never touches N=1024. */
void orc_ref_roundtrip_n(int n, const uint32_t* a, uint32_t* out) {
    void* h = R.mk(n);
    double* f = malloc(sizeof(double) * n);
    R.ifft_u32(h, f, a);
    R.fft_u32(h, out, f);
    free(f);
}
void orc_ref_poly_mul_n(int n, const uint32_t* a, const uint32_t* b, uint32_t* out) { R.poly_mul(R.mk(n), out, a, b); }
void orc_ref_free(void* p) { free(p); }
/* TRGSWRepF::from, hom_nand/src/trgsw.rs:68-76 : struct { cipher_f[2L], pkey_f[2L] }, each ifft_torus of the torus poly */
static void trgsw_to_fourier(const uint32_t* trgsw, double* out) {
    void* h = ref_handle();
    for (int j = 0; j < 2 * L_; j++) {
        R.ifft_u32(h, out + (size_t)(0 * 2 * L_ + j) * N_, trgsw + (size_t)(j * 2 + 0) * N_);
        R.ifft_u32(h, out + (size_t)(1 * 2 * L_ + j) * N_, trgsw + (size_t)(j * 2 + 1) * N_);
    }
}
double* orc_ref_bk_fourier(const uint32_t* bk) {
    size_t per = (size_t)2 * L_ * 2 * N_;
    double* f = malloc(sizeof(double) * per * n_);
    for (int i = 0; i < n_; i++) trgsw_to_fourier(bk + per * i, f + per * i);
    return f;
}
/* FrrSeries::hadamard + add, utils/src/spqlios.rs:161-165,204-222 : layout re[0..N/2) || im[0..N/2) */
static inline void hadamard_acc(double* acc, const double* l, const double* r) {
    const int h = N_ / 2;
    for (int i = 0; i < h; i++) {
        double ii = l[h + i] * r[h + i], rr = l[i] * r[i], ri = l[i] * r[h + i], ir = l[h + i] * r[i];
        acc[i] += rr - ii;
        acc[h + i] += ir + ri;
    }
}
/* TRGSWRepF::cross, hom_nand/src/trgsw.rs:264-306 */
void orc_ref_external_product(const double* gF, const uint32_t* trlwe, uint32_t mask, uint32_t* out) {
    void* h = ref_handle();
    int32_t dig[2 * L_ * N_];
    orc_decompose(trlwe, mask, dig);
    orc_decompose(trlwe + N_, mask, dig + L_ * N_);
    double df[2 * L_][N_];
    for (int j = 0; j < 2 * L_; j++) R.ifft_i32(h, df[j], dig + j * N_);
    double cb[N_], ca[N_];
    memset(cb, 0, sizeof cb); memset(ca, 0, sizeof ca);
    for (int j = 0; j < 2 * L_; j++) hadamard_acc(cb, gF + (size_t)(0 * 2 * L_ + j) * N_, df[j]);
    for (int j = 0; j < 2 * L_; j++) hadamard_acc(ca, gF + (size_t)(1 * 2 * L_ + j) * N_, df[j]);
    R.fft_u32(h, out, cb);
    R.fft_u32(h, out + N_, ca);
}
void orc_ref_external_product_torus(const uint32_t* trgsw, const uint32_t* trlwe, uint32_t mask, uint32_t* out) {
    double* f = malloc(sizeof(double) * 2 * L_ * 2 * N_);
    trgsw_to_fourier(trgsw, f);
    orc_ref_external_product(f, trlwe, mask, out);
    free(f);
}
static void ep_ref(const void* ctx, int i, const uint32_t* trlwe, uint32_t mask, uint32_t* out) {
    size_t per = (size_t)2 * L_ * 2 * N_;
    orc_ref_external_product((const double*)ctx + per * i, trlwe, mask, out);
}
void orc_ref_blind_rotate(const double* bkF, const uint32_t* tlwe, uint32_t mask, int nsteps, uint32_t* out) {
    blind_rotate_generic(ep_ref, bkF, tlwe, mask, nsteps, out);
}
static void ref_bootstrap_one(const double* bkF, const uint32_t* ksk, const uint32_t* in, uint32_t mask, uint32_t* out) {
    uint32_t acc[2 * N_], lwe1[N_ + 1];
    orc_ref_blind_rotate(bkF, in, mask, n_, acc);
    orc_sample_extract0(acc, lwe1);
    orc_key_switch(ksk, lwe1, out);
}
void orc_ref_bootstrap(const double* bkF, const uint32_t* ksk, const uint32_t* in, size_t B, uint32_t mask, int nthreads,
                       uint32_t* out) {
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
    for (long g = 0; g < (long)B; g++) ref_bootstrap_one(bkF, ksk, in + (size_t)g * (n_ + 1), mask, out + (size_t)g * (n_ + 1));
}
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
/* the homnand-bench / hom_nand_profile workload (examples/homnand-bench.rs:22-45, nander/src/lib.rs:187-195):
 * B gates of `op`, each = pre-combination + bootstrap; returns wall seconds. */
double orc_ref_bench_gates(const double* bkF, const uint32_t* ksk, const uint32_t* in0, const uint32_t* in1, int op, size_t B,
                           uint32_t mask, int nthreads, uint32_t* out) {
    if (nthreads < 1) nthreads = 1;
    double t0 = now_s();
#pragma omp parallel for schedule(dynamic) num_threads(nthreads)
    for (long g = 0; g < (long)B; g++) {
        uint32_t lin[n_ + 1];
        orc_gate_linear(op, in0 + (size_t)g * (n_ + 1), in1 ? in1 + (size_t)g * (n_ + 1) : NULL, 1, lin);
        ref_bootstrap_one(bkF, ksk, lin, mask, out + (size_t)g * (n_ + 1));
    }
    return now_s() - t0;
}
