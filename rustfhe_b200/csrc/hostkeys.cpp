// hostkeys.cpp -- host-side key generation, encryption and decryption of the product library.
//
// API surface kept from the reference: TFHE::new = KeySwitchingKey::new + BootstrappingKey::new
// (hom_nand/src/tfhe.rs:21-25,119-126; tlwe.rs:247-277; trgsw.rs:117-139,213-229; trlwe.rs:127-137) and
// Cryptor::{encrypto,decrypto} for TLWE bits (digest.rs:14-33; tlwe.rs:181-240).  One-off, host side, like the reference.
// Deliberate differences: a seeded counter-based generator (the reference cannot be seeded), full 32-bit uniform masks
// (the reference samples Uniform<f32>, 24 random bits, math.rs:425-432) and an exact a*s product (the reference uses its FFT).
#include <cmath>
#include <sys/random.h>
#include <cstring>
#include <thread>
#include <vector>
#include "../../include/tfhe_b200.h"
#include "tfhe_rng.cuh"

namespace {
constexpr int n = 635, N = 1024, L = 3, BGBIT = 6, KS_T = 8, KS_BB = 2;
constexpr uint32_t MU = 0x20000000u;
using namespace tfhe_rng;   // seeded counter-based generator + deterministic Gaussian shared with the device kernels

template <class F>
void parallel_for(int count, F f) {
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    if (nt > 32) nt = 32;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; t++)
        th.emplace_back([=] { for (int i = (int)t; i < count; i += (int)nt) f(i); });
    for (auto& x : th) x.join();
}
// b += a * s mod (X^N + 1) for a binary s given as the list of its set positions
void add_mul_binary(const uint32_t* a, const std::vector<int>& ones, uint32_t* b) {
    for (int j : ones) {
        for (int k = j; k < N; k++) b[k] += a[k - j];
        for (int k = 0; k < j; k++) b[k] -= a[N + k - j];
    }
}
// 32 bytes of operating-system entropy (getrandom(2), blocking until the pool is initialised)
bool os_random(uint8_t* out, size_t len) {
    size_t got = 0;
    while (got < len) {
        const ssize_t r = getrandom(out + got, len - got, 0);
        if (r < 0) return false;
        got += (size_t)r;
    }
    return true;
}
// generator key of a *_csprng call: the caller's 32 bytes, or fresh OS entropy when key == NULL
bool csprng_key(const uint8_t* key, RngKey* out) {
    uint8_t fresh[32];
    if (!key) {
        if (!os_random(fresh, sizeof fresh)) return false;
        key = fresh;
    }
    *out = key_from_bytes(key);
    volatile uint8_t* w = fresh;
    for (size_t i = 0; i < sizeof fresh; i++) w[i] = 0;
    return true;
}
int keygen_secret_impl(const RngKey& seed, uint8_t* s0, uint8_t* s1);
int keygen_bk_impl(const RngKey& seed, const uint8_t* s0, const uint8_t* s1, uint32_t* bk);
int keygen_ksk_impl(const RngKey& seed, const uint8_t* s0, const uint8_t* s1, uint32_t* ksk);
int encrypt_bits_impl(const RngKey& seed, uint64_t ct_index0, const uint8_t* s0, const uint8_t* bits, size_t B, uint32_t* out);
}  // namespace

extern "C" {

// ---- deterministic TEST entry points (SplitMix64 of a 64-bit seed: reproducible, NOT secure; see tfhe_rng.cuh) ----
int tfhe_b200_keygen_secret(uint64_t seed, uint8_t* s0, uint8_t* s1) { return keygen_secret_impl(key_from_seed(seed), s0, s1); }
int tfhe_b200_keygen_bk(uint64_t seed, const uint8_t* s0, const uint8_t* s1, uint32_t* bk) { return keygen_bk_impl(key_from_seed(seed), s0, s1, bk); }
int tfhe_b200_keygen_ksk(uint64_t seed, const uint8_t* s0, const uint8_t* s1, uint32_t* ksk) { return keygen_ksk_impl(key_from_seed(seed), s0, s1, ksk); }
int tfhe_b200_encrypt_bits(uint64_t seed, uint64_t ct_index0, const uint8_t* s0, const uint8_t* bits, size_t B, uint32_t* out) {
    return encrypt_bits_impl(key_from_seed(seed), ct_index0, s0, bits, B, out);
}
// ---- production entry points: ChaCha20 keyed with 256 bits; key == NULL draws a fresh key from getrandom(2) per call ----
int tfhe_b200_random_bytes(uint8_t* out, size_t len) {
    if (!out && len) return TFHE_B200_ERR_PARAM;
    return os_random(out, len) ? TFHE_B200_OK : TFHE_B200_ERR_IO;
}
int tfhe_b200_keygen_secret_csprng(const uint8_t* key, uint8_t* s0, uint8_t* s1) {
    RngKey k;
    if (!csprng_key(key, &k)) return TFHE_B200_ERR_IO;
    return keygen_secret_impl(k, s0, s1);
}
int tfhe_b200_keygen_bk_csprng(const uint8_t* key, const uint8_t* s0, const uint8_t* s1, uint32_t* bk) {
    RngKey k;
    if (!csprng_key(key, &k)) return TFHE_B200_ERR_IO;
    return keygen_bk_impl(k, s0, s1, bk);
}
int tfhe_b200_keygen_ksk_csprng(const uint8_t* key, const uint8_t* s0, const uint8_t* s1, uint32_t* ksk) {
    RngKey k;
    if (!csprng_key(key, &k)) return TFHE_B200_ERR_IO;
    return keygen_ksk_impl(k, s0, s1, ksk);
}
int tfhe_b200_encrypt_bits_csprng(const uint8_t* key, const uint8_t* s0, const uint8_t* bits, size_t B, uint32_t* out) {
    RngKey k;
    if (!csprng_key(key, &k)) return TFHE_B200_ERR_IO;
    return encrypt_bits_impl(k, 0, s0, bits, B, out);
}
}  // extern "C"

namespace {
int keygen_secret_impl(const RngKey& seed, uint8_t* s0, uint8_t* s1) {
    if (!s0 || !s1) return TFHE_B200_ERR_PARAM;
    Rng r0(seed, S0), r1(seed, S1);
    for (int i = 0; i < n; i++) s0[i] = (uint8_t)(r0.u64(i) >> 63);
    for (int i = 0; i < N; i++) s1[i] = (uint8_t)(r1.u64(i) >> 63);
    return TFHE_B200_OK;
}

// BK_i = TRGSW_{s1}(s0_i): 2l fresh TRLWE_{s1}(0) rows (B = A*s1 + e, A), then mu/Bg^(j+1) added on B[0] of rows j<l and on
// A[0] of rows l+j (trgsw.rs:118-138,213-229); alpha_bk = 2^-25 (trlwe.rs:77)
int keygen_bk_impl(const RngKey& seed, const uint8_t* s0, const uint8_t* s1, uint32_t* bk) {
    if (!s0 || !s1 || !bk) return TFHE_B200_ERR_PARAM;
    std::vector<int> ones;
    for (int j = 0; j < N; j++) if (s1[j]) ones.push_back(j);
    const Rng ra(seed, BK_A), re(seed, BK_E);
    parallel_for(n * 2 * L, [&](int row) {
        const int i = row / (2 * L), j = row % (2 * L);
        uint32_t* B = bk + ((size_t)row * 2 + 0) * N;
        uint32_t* A = bk + ((size_t)row * 2 + 1) * N;
        const uint64_t base = (uint64_t)row * N;
        for (int k = 0; k < N; k++) { A[k] = ra.u32(base + k); B[k] = re.gauss(base + k, SCALE_BK); }
        add_mul_binary(A, ones, B);
        const uint32_t mu = (uint32_t)s0[i] << (32 - BGBIT * ((j % L) + 1));
        if (j < L) B[0] += mu; else A[0] += mu;
    });
    return TFHE_B200_OK;
}

// KS[i][l][d-1] = TLWE_{s0}(d * s1_i / 2^(2(l+1))), d = 1..3 (tlwe.rs:247-283; the unreachable d=4 entry is not stored);
// alpha = 2^-15 (tlwe.rs:176)
int keygen_ksk_impl(const RngKey& seed, const uint8_t* s0, const uint8_t* s1, uint32_t* ksk) {
    if (!s0 || !s1 || !ksk) return TFHE_B200_ERR_PARAM;
    const Rng ra(seed, KSK_A), re(seed, KSK_E);
    parallel_for(N * KS_T * 3, [&](int rowid) {
        const int i = rowid / (KS_T * 3), l = (rowid / 3) % KS_T, d = rowid % 3 + 1;
        uint32_t* row = ksk + (size_t)rowid * (n + 1);
        uint32_t b = ((uint32_t)(d * s1[i]) << (32 - KS_BB * (l + 1))) + re.gauss((uint64_t)rowid, SCALE_LV0);
        for (int c = 0; c < n; c++) {
            const uint32_t a = ra.u32((uint64_t)rowid * n + c);
            row[1 + c] = a;
            if (s0[c]) b += a;
        }
        row[0] = b;
    });
    return TFHE_B200_OK;
}

// Cryptor::encrypto(TLWE, &s0, Binary): One -> +1/8, Zero -> -1/8 (tlwe.rs:181-186), b = <a,s> + e + m (tlwe.rs:213-228)
int encrypt_bits_impl(const RngKey& seed, uint64_t ct_index0, const uint8_t* s0, const uint8_t* bits, size_t B, uint32_t* out) {
    if (!s0 || (!bits && B) || (!out && B)) return TFHE_B200_ERR_PARAM;
    const Rng ra(seed, ENC_A), re(seed, ENC_E);
    for (size_t g = 0; g < B; g++) {
        uint32_t* ct = out + g * (n + 1);
        const uint64_t id = ct_index0 + g;
        uint32_t b = (bits[g] ? MU : 0u - MU) + re.gauss(id, SCALE_LV0);
        for (int i = 0; i < n; i++) {
            const uint32_t a = ra.u32(id * n + i);
            ct[1 + i] = a;
            if (s0[i]) b += a;
        }
        ct[0] = b;
    }
    return TFHE_B200_OK;
}
}  // namespace

extern "C" {
// phase = b - <a, s> (tlwe.rs:230-240)
int tfhe_b200_phase(const uint8_t* s0, const uint32_t* ct, size_t B, uint32_t* phase) {
    if (!s0 || (!ct && B) || (!phase && B)) return TFHE_B200_ERR_PARAM;
    for (size_t g = 0; g < B; g++) {
        const uint32_t* c = ct + g * (n + 1);
        uint32_t acc = c[0];
        for (int i = 0; i < n; i++) if (s0[i]) acc -= c[1 + i];
        phase[g] = acc;
    }
    return TFHE_B200_OK;
}
// torus2binary: f32(phase) < 0.5 -> One (tlwe.rs:187-194, math.rs:684-690)
int tfhe_b200_decrypt_bits(const uint8_t* s0, const uint32_t* ct, size_t B, uint8_t* bits) {
    if (!s0 || (!ct && B) || (!bits && B)) return TFHE_B200_ERR_PARAM;
    for (size_t g = 0; g < B; g++) {
        uint32_t ph;
        tfhe_b200_phase(s0, ct + g * (n + 1), 1, &ph);
        bits[g] = ((float)ph * (1.0f / 4294967296.0f)) < 0.5f ? 1 : 0;
    }
    return TFHE_B200_OK;
}
}
