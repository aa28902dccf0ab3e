// homnand_bench.cpp -- the reference's examples/homnand-bench.rs through the C++ host side (include/tfhe_b200.hpp):
// truth tables of nand / and / or / xor / not / mux on fresh encryptions, every gate timed on its own, then the same gates
// as batches.  Exit code 0 = every decryption is right; 2 = a wrong bit; 3 = no engine (e.g. no B200: there is no fallback).
//
//   g++ -O2 -std=c++17 -Iinclude examples/homnand_bench.cpp -Lrustfhe_b200 -lrustfhe_b200 -Wl,-rpath,$PWD/rustfhe_b200 -o homnand_bench
#include <chrono>
#include <cstdio>
#include <functional>
#include <memory>
#include <string>
#include "tfhe_b200.hpp"

using namespace tfhe;
using Clock = std::chrono::steady_clock;

static int failures = 0;

template <class F>
static auto timeit(const std::string& title, F&& f) {   // utils::timeit! (utils/src/lib.rs) prints the wall time of the expression
    const auto t0 = Clock::now();
    auto r = f();
    const double us = std::chrono::duration<double, std::micro>(Clock::now() - t0).count();
    std::printf("%-12s %10.1f us\n", title.c_str(), us);
    return r;
}

int main(int argc, char** argv) {
    // no argument: keys and encryptions from the OS-keyed ChaCha20 generator, like the reference's thread_rng;
    // a seed argument selects the deterministic TEST generator (reproducible runs, NOT secure)
    const bool seeded = argc > 1;
    const uint64_t seed = seeded ? std::stoull(argv[1], nullptr, 0) : 0;
    uint64_t next_index = 0;
    try {
        SecretKeyLv0 s_key_tlwelv0;
        SecretKeyLv1 s_key_tlwelv1;
        if (seeded) gen_secret_keys(seed, s_key_tlwelv0, s_key_tlwelv1); else gen_secret_keys(s_key_tlwelv0, s_key_tlwelv1);
        const auto t0 = Clock::now();
        std::unique_ptr<TFHE> tfhe_p(seeded ? new TFHE(s_key_tlwelv0, s_key_tlwelv1, TFHE::TestSeed{seed}) : new TFHE(s_key_tlwelv0, s_key_tlwelv1));
        TFHE& tfhe = *tfhe_p;
        std::printf("%s\nTFHE::new %.1f ms (both keys generated on the device)\n", tfhe_b200_version(),
                    std::chrono::duration<double, std::milli>(Clock::now() - t0).count());
        auto tlwelv0_ = [&](Binary b) {
            return seeded ? Cryptor::encrypto(TLWE, s_key_tlwelv0, b, seed, next_index++) : Cryptor::encrypto(TLWE, s_key_tlwelv0, b);
        };
        auto dec = [&](const TLWERep& r) { return Cryptor::decrypto(TLWE, s_key_tlwelv0, r); };
        auto bin = [](int v) { return v ? Binary::One : Binary::Zero; };

        struct Gate2 { const char* title; std::function<TLWERep(const TLWERep&, const TLWERep&)> f; int truth[4]; };
        const Gate2 gates[] = {
            {"nand", [&](const TLWERep& a, const TLWERep& b) { return tfhe.hom_nand(a, b); }, {1, 1, 1, 0}},
            {"and", [&](const TLWERep& a, const TLWERep& b) { return tfhe.hom_and(a, b); }, {0, 0, 0, 1}},
            {"or", [&](const TLWERep& a, const TLWERep& b) { return tfhe.hom_or(a, b); }, {0, 1, 1, 1}},
            {"xor", [&](const TLWERep& a, const TLWERep& b) { return tfhe.hom_xor(a, b); }, {0, 1, 1, 0}},
        };
        for (const Gate2& g : gates)
            for (int i = 0; i < 4; i++) {   // input_0 = bit 0 of i, input_1 = bit 1 of i (homnand-bench.rs:25-27)
                const TLWERep in0 = tlwelv0_(bin(i & 1)), in1 = tlwelv0_(bin(i & 2));
                const TLWERep rep = timeit(std::string(g.title) + " " + std::to_string(i & 1) + " " + std::to_string((i >> 1) & 1), [&] { return g.f(in0, in1); });
                if (dec(rep) != bin(g.truth[i])) { std::printf("  WRONG: %s %d %d\n", g.title, i & 1, (i >> 1) & 1); failures++; }
            }
        for (int i = 0; i < 2; i++) {
            const TLWERep in = tlwelv0_(bin(i));
            const TLWERep rep = timeit("not " + std::to_string(i), [&] { return tfhe.hom_not(in); });
            if (dec(rep) != bin(!i)) { std::printf("  WRONG: not %d\n", i); failures++; }
        }
        for (int i = 0; i < 8; i++) {       // hom_mux(control, input_0, input_1) = control ? input_1 : input_0 (tfhe.rs:27-40)
            const int c = i & 1, x0 = (i >> 1) & 1, x1 = (i >> 2) & 1;
            const TLWERep rep = timeit("mux " + std::to_string(c) + " " + std::to_string(x0) + " " + std::to_string(x1),
                                       [&] { return tfhe.hom_mux(tlwelv0_(bin(c)), tlwelv0_(bin(x0)), tlwelv0_(bin(x1))); });
            if (dec(rep) != bin(c ? x1 : x0)) { std::printf("  WRONG: mux %d %d %d\n", c, x0, x1); failures++; }
        }
        // the same through the batch form: 1024 independent NAND gates in one call
        const size_t B = 1024;
        std::vector<TLWERep> a(B), b(B);
        for (size_t k = 0; k < B; k++) { a[k] = tlwelv0_(bin((k * 7 + 1) & 4)); b[k] = tlwelv0_(bin((k * 5 + 3) & 8)); }
        tfhe.hom_nand_batch(a, b);   // warm-up (workspace allocation)
        const auto t1 = Clock::now();
        const std::vector<TLWERep> out = tfhe.hom_nand_batch(a, b);
        const double ms = std::chrono::duration<double, std::milli>(Clock::now() - t1).count();
        for (size_t k = 0; k < B; k++)
            if (dec(out[k]) != bin(!(((k * 7 + 1) & 4) && ((k * 5 + 3) & 8)))) failures++;
        std::printf("hom_nand_batch: %zu gates in %.2f ms = %.0f gates/s (host buffers, copies included)\n", B, ms, B / ms * 1e3);
    } catch (const Error& e) {
        std::fprintf(stderr, "tfhe::Error %d: %s\n", e.code, e.what());
        return 3;
    }
    std::printf(failures ? "FAILED: %d wrong decryptions\n" : "all decryptions right\n", failures);
    return failures ? 2 : 0;
}
