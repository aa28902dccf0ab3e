// tfhe_b200.hpp -- C++ host side above the C ABI (tfhe_b200.h), mirroring the reference crate `hom_nand` for the gate path:
// same names, same argument meaning, same results.  Header only; link librustfhe_b200.
//
//   reference (Rust)                                             here
//   ------------------------------------------------------------ ----------------------------------------------------------
//   utils::math::Binary                         (math.rs:561-606) tfhe::Binary { Zero, One }
//   tlwe::TLWEHelper::N = 635, tfhe::TFHEHelper::NBIT = 10        tfhe::TLWEHelper::N, tfhe::TFHEHelper::{NBIT, N}
//   tlwe::TLWERep<N>  { cipher, p_key }         (tlwe.rs:19-41)  tfhe::TLWERep  ([b, a_0..a_{n-1}] words, the ABI layout)
//     TLWERep::trivial, AsLogic                 (tlwe.rs:76-87)    TLWERep::trivial, logic_true, logic_false
//   digest::Cryptor::{encrypto, decrypto}(TLWE) (digest.rs:14-33) tfhe::Cryptor::{encrypto, decrypto}(TLWE, ...)
//   tfhe::TFHE::new(s_key_tlwelv0, s_key_tlwelv1) (tfhe.rs:21-25) tfhe::TFHE(s_key_tlwelv0, s_key_tlwelv1, seed, device)
//   TFHE::hom_nand / and / or / xor / not / mux (tfhe.rs:27-71)   TFHE::hom_nand / ... / hom_mux  (one gate: a batch of one)
//   --                                                            TFHE::hom_*_batch(vector<TLWERep>, ...)  one launch per batch
//
// Randomness: like the reference (rand::thread_rng, an OS-seeded ChaCha generator) keys and encryptions come from ChaCha20
// keyed by getrandom(2) -- the *_csprng entry points of the C ABI.  The overloads that take a `seed` use the DETERMINISTIC
// TEST generator instead (reproducible, NOT secure: parity tests only).  Differences, all deliberate: errors are C++ exceptions (tfhe::Error carries the TFHE_B200_ERR_* code and the engine's message) where
// the reference aborts; there is no CPU fallback -- constructing a TFHE without a B200 throws.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include "tfhe_b200.h"

namespace tfhe {

enum class Binary : uint8_t { Zero = 0, One = 1 };   // math.rs:561-606
inline Binary operator!(Binary b) { return b == Binary::One ? Binary::Zero : Binary::One; }

struct TLWEHelper { static constexpr int N = 635; };                      // tlwe.rs:175
struct TFHEHelper { static constexpr int NBIT = 10, N = 1 << NBIT; };      // tfhe.rs:14-19
struct TLWEStrategy {};                                                    // the `TLWE` unit struct of tlwe.rs:172
inline constexpr TLWEStrategy TLWE{};

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};
inline void check(int rc, const tfhe_b200_ctx* ctx, const char* where) {
    if (rc == TFHE_B200_OK) return;
    const char* msg = tfhe_b200_last_error(ctx);
    throw Error(rc, std::string(where) + ": " + (msg && *msg ? msg : "error") + " (code " + std::to_string(rc) + ")");
}

using SecretKeyLv0 = std::array<Binary, TLWEHelper::N>;
using SecretKeyLv1 = std::array<Binary, TFHEHelper::N>;

// uniform secret keys (the reference draws them with BinaryDistribution::uniform(), homnand-bench.rs:10-12): OS-keyed ChaCha20
inline void gen_secret_keys(SecretKeyLv0& s0, SecretKeyLv1& s1) {
    static_assert(sizeof(Binary) == 1, "Binary is one byte: the arrays are the ABI's bit arrays");
    check(tfhe_b200_keygen_secret_csprng(nullptr, reinterpret_cast<uint8_t*>(s0.data()), reinterpret_cast<uint8_t*>(s1.data())), nullptr, "keygen_secret");
}
// deterministic TEST form (NOT secure)
inline void gen_secret_keys(uint64_t seed, SecretKeyLv0& s0, SecretKeyLv1& s1) {
    check(tfhe_b200_keygen_secret(seed, reinterpret_cast<uint8_t*>(s0.data()), reinterpret_cast<uint8_t*>(s1.data())), nullptr, "keygen_secret");
}

struct TLWERep {                                        // tlwe.rs:19-41
    std::array<uint32_t, TLWEHelper::N + 1> w{};        // w[0] = cipher (b), w[1..] = p_key (a)
    uint32_t cipher() const { return w[0]; }
    const uint32_t* p_key() const { return w.data() + 1; }
    static TLWERep trivial(uint32_t text) { TLWERep r; r.w[0] = text; return r; }   // tlwe.rs:76-79
    static TLWERep logic_true() { return trivial(0x20000000u); }                    // AsLogic, tlwe.rs:80-87
    static TLWERep logic_false() { return trivial(0xE0000000u); }
};

struct Cryptor {                                        // digest.rs:14-33 with the TLWE strategy of tlwe.rs:197-241
    // fresh encryption of one bit under s_key: mask and noise from ChaCha20 with a fresh getrandom(2) key per call
    static TLWERep encrypto(TLWEStrategy, const SecretKeyLv0& s_key, Binary item) {
        TLWERep r;
        const uint8_t bit = (uint8_t)item;
        check(tfhe_b200_encrypt_bits_csprng(nullptr, reinterpret_cast<const uint8_t*>(s_key.data()), &bit, 1, r.w.data()), nullptr, "encrypt_bits");
        return r;
    }
    // deterministic TEST form (NOT secure): (seed, index) select the mask and the noise; the caller keeps the pairs distinct
    static TLWERep encrypto(TLWEStrategy, const SecretKeyLv0& s_key, Binary item, uint64_t seed, uint64_t index) {
        TLWERep r;
        const uint8_t bit = (uint8_t)item;
        check(tfhe_b200_encrypt_bits(seed, index, reinterpret_cast<const uint8_t*>(s_key.data()), &bit, 1, r.w.data()), nullptr, "encrypt_bits");
        return r;
    }
    static Binary decrypto(TLWEStrategy, const SecretKeyLv0& s_key, const TLWERep& rep) {
        uint8_t bit = 0;
        check(tfhe_b200_decrypt_bits(reinterpret_cast<const uint8_t*>(s_key.data()), rep.w.data(), 1, &bit), nullptr, "decrypt_bits");
        return bit ? Binary::One : Binary::Zero;
    }
};

class TFHE {                                            // tfhe.rs:9-113
public:
    // TFHE::new: both evaluation keys are generated ON the device, masks and noise from ChaCha20 keyed by getrandom(2)
    TFHE(const SecretKeyLv0& s_key_tlwelv0, const SecretKeyLv1& s_key_tlwelv1, int device = 0) { init(s_key_tlwelv0, s_key_tlwelv1, nullptr, device); }
    // deterministic TEST form (NOT secure): bit-identical to the seeded host keygen
    struct TestSeed { uint64_t seed; };
    TFHE(const SecretKeyLv0& s_key_tlwelv0, const SecretKeyLv1& s_key_tlwelv1, TestSeed seed, int device = 0) { init(s_key_tlwelv0, s_key_tlwelv1, &seed, device); }
    ~TFHE() { if (ctx_) tfhe_b200_ctx_destroy(ctx_); }
    TFHE(const TFHE&) = delete;
    TFHE& operator=(const TFHE&) = delete;

private:
    void init(const SecretKeyLv0& s_key_tlwelv0, const SecretKeyLv1& s_key_tlwelv1, const TestSeed* seed, int device) {
        check(tfhe_b200_ctx_create(nullptr, device, &ctx_), nullptr, "ctx_create");
        const uint8_t* k0 = reinterpret_cast<const uint8_t*>(s_key_tlwelv0.data());
        const uint8_t* k1 = reinterpret_cast<const uint8_t*>(s_key_tlwelv1.data());
        const int rc = seed ? tfhe_b200_keygen_device(ctx_, seed->seed, k0, k1) : tfhe_b200_keygen_device_csprng(ctx_, nullptr, k0, k1);
        if (rc != TFHE_B200_OK) {
            const std::string msg = tfhe_b200_last_error(ctx_);
            tfhe_b200_ctx_destroy(ctx_);
            throw Error(rc, "keygen_device: " + msg);
        }
    }

public:

    TLWERep hom_nand(const TLWERep& input_0, const TLWERep& input_1) const { return gate(TFHE_B200_NAND, input_0, &input_1); }   // tfhe.rs:41-47
    TLWERep hom_and(const TLWERep& input_0, const TLWERep& input_1) const { return gate(TFHE_B200_AND, input_0, &input_1); }     // tfhe.rs:48-54
    TLWERep hom_or(const TLWERep& input_0, const TLWERep& input_1) const { return gate(TFHE_B200_OR, input_0, &input_1); }       // tfhe.rs:55-61
    TLWERep hom_xor(const TLWERep& input_0, const TLWERep& input_1) const { return gate(TFHE_B200_XOR, input_0, &input_1); }     // tfhe.rs:62-68
    TLWERep hom_not(const TLWERep& input) const { return gate(TFHE_B200_NOT, input, nullptr); }                                  // tfhe.rs:69-71
    // control ? input_1 : input_0 (tfhe.rs:27-40)
    TLWERep hom_mux(const TLWERep& control, const TLWERep& input_0, const TLWERep& input_1) const {
        TLWERep out;
        check(tfhe_b200_mux_batch(ctx_, control.w.data(), input_0.w.data(), input_1.w.data(), out.w.data(), 1), ctx_, "mux_batch");
        return out;
    }
    // the batch forms: B independent gates in one launch pair (what the engine is built for)
    std::vector<TLWERep> hom_nand_batch(const std::vector<TLWERep>& a, const std::vector<TLWERep>& b) const { return gates(TFHE_B200_NAND, a, &b); }
    std::vector<TLWERep> hom_and_batch(const std::vector<TLWERep>& a, const std::vector<TLWERep>& b) const { return gates(TFHE_B200_AND, a, &b); }
    std::vector<TLWERep> hom_or_batch(const std::vector<TLWERep>& a, const std::vector<TLWERep>& b) const { return gates(TFHE_B200_OR, a, &b); }
    std::vector<TLWERep> hom_xor_batch(const std::vector<TLWERep>& a, const std::vector<TLWERep>& b) const { return gates(TFHE_B200_XOR, a, &b); }
    std::vector<TLWERep> hom_not_batch(const std::vector<TLWERep>& a) const { return gates(TFHE_B200_NOT, a, nullptr); }
    std::vector<TLWERep> hom_mux_batch(const std::vector<TLWERep>& control, const std::vector<TLWERep>& input_0, const std::vector<TLWERep>& input_1) const {
        if (control.size() != input_0.size() || control.size() != input_1.size()) throw Error(TFHE_B200_ERR_PARAM, "hom_mux_batch: operand batches differ in length");
        std::vector<TLWERep> out(control.size());
        if (control.empty()) return out;
        if (2 * control.size() > reserved_) reserve(2 * control.size());   // the first stage runs both AND batches in one launch
        check(tfhe_b200_mux_batch(ctx_, control[0].w.data(), input_0[0].w.data(), input_1[0].w.data(), out[0].w.data(), control.size()), ctx_, "mux_batch");
        return out;
    }

    // pre-allocates every work slot of the engine for batches of up to max_batch gates (done automatically by the batch forms
    // the first time a larger batch arrives: no cudaMalloc in later calls)
    void reserve(size_t max_batch) const {
        check(tfhe_b200_reserve(ctx_, max_batch), ctx_, "reserve");
        if (max_batch > reserved_) reserved_ = max_batch;
    }
    tfhe_b200_ctx* raw() const { return ctx_; }

private:
    static_assert(sizeof(TLWERep) == (TLWEHelper::N + 1) * sizeof(uint32_t), "a vector<TLWERep> is the ABI's [B][n+1] array");
    TLWERep gate(int op, const TLWERep& a, const TLWERep* b) const {
        TLWERep out;
        check(tfhe_b200_gate_batch(ctx_, op, a.w.data(), b ? b->w.data() : nullptr, out.w.data(), 1), ctx_, "gate_batch");
        return out;
    }
    std::vector<TLWERep> gates(int op, const std::vector<TLWERep>& a, const std::vector<TLWERep>* b) const {
        if (b && b->size() != a.size()) throw Error(TFHE_B200_ERR_PARAM, "hom_*_batch: operand batches differ in length");
        std::vector<TLWERep> out(a.size());
        if (a.empty()) return out;
        if (a.size() > reserved_) reserve(a.size());
        check(tfhe_b200_gate_batch(ctx_, op, a[0].w.data(), b ? (*b)[0].w.data() : nullptr, out[0].w.data(), a.size()), ctx_, "gate_batch");
        return out;
    }
    tfhe_b200_ctx* ctx_ = nullptr;
    mutable size_t reserved_ = 0;
};

}  // namespace tfhe
