/*
 * tfhe_oracle.h -- CPU ORACLE for the bootstrapped-HomNAND path of hideki1217/rusTfhe.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (librustfhe_b200.so)
 * never links, loads or calls anything in oracle/.
 *
 * Two layers:
 *   exact layer     : exact integer arithmetic mod 2^32 (schoolbook + an independent 2-prime CPU NTT).
 *                     Ground truth the CUDA kernels must match BIT-EXACTLY.
 *   reference layer : the reference's gate path restated in C, calling the reference's OWN native FFT
 *                     (oracle/_ref/libspqlios_ref.so, compiled in place from /root/reference/utils/src/spqlios).
 *                     The CUDA path must match it on decrypted bits and within the stated phase bound.
 *
 * Parity pinning: the reference holds no ciphertext-level golden vectors and no seeding API,
 * so this oracle is pinned by (a) every arithmetic KAT the reference's tests hold for the path
 * (tests/test_oracle_kats.py) and (b) outputs of the reference's real FFT library run in-process.
 *
 * All citations are file:line relative to /root/reference.
 */
#ifndef TFHE_ORACLE_H
#define TFHE_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* default parameter set: tlwe.rs:175-180, trlwe.rs:76-77, trgsw.rs:112-115, tfhe.rs:16-17 */
#define ORC_n 635
#define ORC_N 1024
#define ORC_L 3
#define ORC_BGBIT 6
#define ORC_KS_T 8
#define ORC_KS_BASEBIT 2
#define ORC_MU 0x20000000u             /* 1/8, tlwe.rs:181-186 */
#define ORC_MASK_FAITHFUL 0x02084000u  /* make_decomp_mask(3,6) as evaluated, math.rs:542-560 */
#define ORC_MASK_TESTED 0x02082000u    /* decomposition_i32's OR-built mask, math.rs:582-591 */

enum { ORC_NAND = 0, ORC_AND = 1, ORC_OR = 2, ORC_XOR = 3, ORC_NOT = 4, ORC_COPY = 5, ORC_ANDNY = 6 };

/* ---- seeded data generation (the harness owns the RNG: SURVEY F3) ---- */
uint64_t orc_rnd64(uint64_t seed, uint64_t stream, uint64_t idx);
int32_t orc_gauss_torus(uint64_t seed, uint64_t stream, uint64_t idx, double alpha);
void orc_keygen_secret(uint64_t seed, uint8_t* s0, uint8_t* s1);
void orc_keygen_bk(uint64_t seed, const uint8_t* s0, const uint8_t* s1, uint32_t* bk /*[n][2L][2][N]*/);
void orc_keygen_ksk(uint64_t seed, const uint8_t* s0, const uint8_t* s1, uint32_t* ksk /*[N][T][3][n+1]*/);
void orc_tlwe_encrypt_bits(uint64_t seed, uint64_t ct_index0, const uint8_t* s0, const uint8_t* bits, size_t B,
                           uint32_t* out /*[B][n+1]*/);
void orc_tlwe_phase(const uint8_t* s0, const uint32_t* ct, size_t B, uint32_t* phase);
void orc_tlwe_decrypt_bits(const uint8_t* s0, const uint32_t* ct, size_t B, uint8_t* bits);
void orc_tlwe1_phase(const uint8_t* s1, const uint32_t* ct /*[B][N+1]*/, size_t B, uint32_t* phase);
void orc_trlwe_phase(const uint8_t* s1, const uint32_t* trlwe /*[2][N]*/, uint32_t* phase /*[N]*/);

/* ---- integer glue on the path (generic n where the reference's KATs need it) ---- */
void orc_rotate(const uint32_t* p, int n_coef, int k, uint32_t* out);
void orc_negacyclic_mul_schoolbook(const uint32_t* a, const int32_t* d, int n_coef, uint32_t* out);
uint32_t orc_make_decomp_mask(uint32_t l, uint32_t bits);
uint32_t orc_tested_decomp_mask(uint32_t l, uint32_t bits);
void orc_decompose_scalar(uint32_t x, uint32_t l, uint32_t bits, uint32_t mask, int32_t* out /*[l]*/);
void orc_decompose(const uint32_t* p, uint32_t mask, int32_t* out /*[L][N]*/);
void orc_decompose_u32_scalar(uint32_t x, uint32_t l, uint32_t bits, uint32_t* out /*[l]*/);
uint32_t orc_torus_from_f32(float v);
float orc_torus_to_f32(uint32_t t);
void orc_gate_linear(int op, const uint32_t* in0, const uint32_t* in1, size_t B, uint32_t* out);
void orc_sample_extract0(const uint32_t* trlwe /*[2][N]*/, uint32_t* out /*[N+1]: b, a*/);
void orc_sample_extract(const uint32_t* trlwe, int index, uint32_t* out);
void orc_key_switch(const uint32_t* ksk, const uint32_t* lwe1 /*[N+1]*/, uint32_t* out /*[n+1]*/);
void orc_ks_digits(const uint32_t* lwe1, uint16_t* dig /*[N] packed 8x2 bit, level 0 in bits 15:14*/);

/* ---- exact layer ---- */
void orc_negacyclic_mul_ntt(const uint32_t* a, const int32_t* d, uint32_t* out); /* N=1024, exact mod 2^32 */
void orc_external_product_exact(const uint32_t* trgsw /*[2L][2][N]*/, const uint32_t* trlwe /*[2][N]*/,
                                uint32_t mask, uint32_t* out /*[2][N]*/);
void* orc_exact_bk_prepare(const uint32_t* bk);
void orc_exact_bk_free(void* h);
void orc_blind_rotate_exact(const void* bkh, const uint32_t* tlwe /*[n+1]*/, uint32_t mask, int nsteps,
                            uint32_t* out_trlwe /*[2][N]*/);
void orc_bootstrap_exact(const void* bkh, const uint32_t* ksk, const uint32_t* in /*[B][n+1]*/, size_t B, uint32_t mask,
                         uint32_t* out /*[B][n+1]*/, uint32_t* out_lwe1 /*[B][N+1] or NULL*/);

/* ---- reference layer (needs oracle/_ref/libspqlios_ref.so) ---- */
int orc_ref_init(const char* libpath); /* 0 ok */
int orc_ref_available(void);
void orc_ref_poly_mul(const uint32_t* a, const uint32_t* b, uint32_t* out); /* Spqlios_poly_mul */
void orc_ref_ifft_fft_roundtrip(const uint32_t* a, uint32_t* out);
double* orc_ref_bk_fourier(const uint32_t* bk); /* [n][cipher_f[2L], pkey_f[2L]][N] doubles; free with orc_ref_free */
void orc_ref_free(void* p);
void orc_ref_roundtrip_n(int n, const uint32_t* a, uint32_t* out); /* separate process only (F6) */
void orc_ref_poly_mul_n(int n, const uint32_t* a, const uint32_t* b, uint32_t* out);
void orc_ref_external_product(const double* trgswF /*[2][2L][N]*/, const uint32_t* trlwe, uint32_t mask, uint32_t* out);
void orc_ref_external_product_torus(const uint32_t* trgsw, const uint32_t* trlwe, uint32_t mask, uint32_t* out);
void orc_ref_blind_rotate(const double* bkF, const uint32_t* tlwe, uint32_t mask, int nsteps, uint32_t* out_trlwe);
void orc_ref_bootstrap(const double* bkF, const uint32_t* ksk, const uint32_t* in, size_t B, uint32_t mask, int nthreads,
                       uint32_t* out);
/* times `reps` bootstrapped gates of op on each of nthreads threads; returns wall seconds */
double orc_ref_bench_gates(const double* bkF, const uint32_t* ksk, const uint32_t* in0, const uint32_t* in1, int op,
                           size_t B, uint32_t mask, int nthreads, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif
