// host_emul.cpp -- CPU execution of the SAME per-lane arithmetic the CUDA kernels run (ntt32.cuh / cmux_steps.cuh),
// lane by lane, with the same shared-memory tiles, swizzles and device key layout.  Built as libhostemul.so and used
// only by the CPU test-suite (tests/test_host_emul.py) to pin the kernel math against the oracle without a GPU.
// It is NOT a fallback: the product C-ABI never calls it.
#include <vector_types.h>
#include <vector_functions.h>
#include <cstring>
#include <vector>
#include "t2_steps.cuh"
#include "fft64.cuh"
#include "tfhe_rng.cuh"

using namespace tfhe;

extern "C" {

// TRGSW torus polys [row j][poly][1024] (reference order: rows b-first, poly 0 = cipher, 1 = p_key) -> device layout
void emul_key_transform(const uint32_t* trgsw, uint32_t* dev) {
    std::vector<uint32_t> S(TILE_WORDS);
    for (int j = 0; j < BK_ROWS; j++)
        for (int poly = 0; poly < 2; poly++)
            for (int part = 0; part < 3; part++) {
                const uint32_t* src = trgsw + (size_t)(j * 2 + poly) * 1024;
                for (int lane = 0; lane < 32; lane++) key_cols(lane, src, part, S.data());
                for (int lane = 0; lane < 32; lane++) key_rows(lane, S.data(), h_fwdB, dev + bk_off(0, poly, part, j, 0, 0));
            }
}

static void cmux_core(const uint32_t* dev, const uint32_t* acc, bool rotate, uint32_t abar, uint32_t mask, uint32_t* sp) {
    std::vector<uint32_t> dh(6 * TILE_WORDS), U(2 * 1024);
    for (int w = 0; w < 6; w++) {   // phase 1u: every warp contributes its rows of the masked source polynomial
        const int poly = w / 3, k = w % 3;
        for (int lane = 0; lane < 32; lane++) {
            if (rotate) p1u<true>(lane, acc + poly * 1024, abar, mask, k, U.data() + poly * 1024);
            else p1u<false>(lane, acc + poly * 1024, abar, mask, k, U.data() + poly * 1024);
        }
    }
    for (int w = 0; w < 6; w++) {
        const int poly = w / 3, k = w % 3;
        uint32_t* S = dh.data() + w * TILE_WORDS;
        for (int lane = 0; lane < 32; lane++) p1a(lane, U.data() + poly * 1024, k, S, h_digit_tab.v);
        for (int lane = 0; lane < 32; lane++) p1b(lane, S, h_fwdB);
    }
    for (int w = 0; w < 6; w++) {
        const int poly = w / 3, k = w % 3;
        uint32_t* S = sp + w * TILE_WORDS;
        uint32_t x[32][32];
        for (int lane = 0; lane < 32; lane++) p2a(lane, dev + bk_off(0, poly, k, 0, 0, 0), dh.data(), h_invB, S);
        for (int lane = 0; lane < 32; lane++) p2b(lane, S, k, x[lane]);
        for (int lane = 0; lane < 32; lane++) p2c(lane, S, x[lane]);
    }
}
// plain external product: out = TRGSW (x) trlwe  (hom_nand/src/trgsw.rs:264-306)
void emul_external_product(const uint32_t* dev, const uint32_t* trlwe, uint32_t mask, uint32_t* out) {
    std::vector<uint32_t> sp(6 * TILE_WORDS);
    cmux_core(dev, trlwe, false, 0, mask, sp.data());
    for (int poly = 0; poly < 2; poly++)
        for (int k = 0; k < 1024; k++)
            out[poly * 1024 + k] = sp[(3 * poly) * TILE_WORDS + k] + sp[(3 * poly + 1) * TILE_WORDS + k] + sp[(3 * poly + 2) * TILE_WORDS + k];
}
// one blind-rotation step: acc <- BK (x) (X^abar acc - acc) + acc   (hom_nand/src/tfhe.rs:103-110)
void emul_cmux_rotate(const uint32_t* dev, uint32_t* acc, uint32_t abar, uint32_t mask) {
    std::vector<uint32_t> sp(6 * TILE_WORDS);
    cmux_core(dev, acc, true, abar, mask, sp.data());
    for (int poly = 0; poly < 2; poly++)
        for (int k = 0; k < 1024; k++)
            acc[poly * 1024 + k] += sp[(3 * poly) * TILE_WORDS + k] + sp[(3 * poly + 1) * TILE_WORDS + k] + sp[(3 * poly + 2) * TILE_WORDS + k];
}
// The same external product through the step functions of the LATENCY kernel (blind_rotate_pair_kernel): both passes of a
// transform in one code body (fwd_shared / inv_shared), the multiply-accumulate split in two halves of the key rows.
// These functions contain a warp-level synchronisation (the transpose through a shared tile), so the 32 lanes of a warp
// are run twice: the first round fills the tile (its own results are discarded), the second reads complete rows / columns;
// what a lane writes does not depend on the round.
void emul_external_product_shared(const uint32_t* dev, const uint32_t* trlwe, uint32_t mask, uint32_t* out) {
    std::vector<uint32_t> dh(6 * TILE_WORDS), U(2 * 1024), S(TILE_WORDS), T(TILE_WORDS);
    for (int w = 0; w < 6; w++)
        for (int lane = 0; lane < 32; lane++) p1u<false>(lane, trlwe + (w / 3) * 1024, 0, mask, w % 3, U.data() + (w / 3) * 1024);
    for (int w = 0; w < 6; w++) {
        uint32_t x[32][32];
        for (int round = 0; round < 2; round++)
            for (int lane = 0; lane < 32; lane++) fwd_shared(lane, U.data() + (w / 3) * 1024, w % 3, S.data(), h_fwdA, h_fwdB, x[lane]);
        for (int lane = 0; lane < 32; lane++)
            for (int q = 0; q < 8; q++)
                *reinterpret_cast<uint4*>(dh.data() + w * TILE_WORDS + swz_chunk(lane, q)) = make_uint4(x[lane][4 * q], x[lane][4 * q + 1], x[lane][4 * q + 2], x[lane][4 * q + 3]);
    }
    memset(out, 0, 2 * 1024 * sizeof(uint32_t));
    for (int w = 0; w < 6; w++) {
        const int poly = w / 3, k = w % 3;
        const uint32_t* slab = dev + bk_off(0, poly, k, 0, 0, 0);
        uint32_t x[32][32];
        for (int round = 0; round < 2; round++)
            for (int lane = 0; lane < 32; lane++) {
                uint64_t mac[32];
                p2a_mac_part(lane, slab, dh.data(), 0, mac, true);
                p2a_mac_part(lane, slab, dh.data() + 3 * TILE_WORDS, 3, mac, false);
                p2a_mac_redc(mac, x[lane]);
                inv_shared(lane, x[lane], T.data(), h_invB, h_invA, k, 3);
            }
        for (int lane = 0; lane < 32; lane++)
            for (int r = 0; r < 32; r++) out[poly * 1024 + 32 * r + lane] += x[lane][r];
    }
}

// ---- throughput kernel (blind_rotate_t2.cuh, t2_steps.cuh): one gate on two warps, two 16-bit key slices ----
// TRGSW torus polys [row j][poly][1024] -> throughput layout [poly][q][j][slice][lane][4] of one step
void emul_key_transform_t2(const uint32_t* trgsw, uint32_t* dev) {
    std::vector<uint32_t> S(TILE_WORDS);
    for (int j = 0; j < BK_ROWS; j++)
        for (int poly = 0; poly < 2; poly++)
            for (int part = 0; part < 2; part++) {
                const uint32_t* src = trgsw + (size_t)(j * 2 + poly) * 1024;
                for (int lane = 0; lane < 32; lane++) key_cols(lane, src, part, S.data(), 2);
                for (int lane = 0; lane < 32; lane++) key_rows_t2(lane, S.data(), h_fwdB, dev + t2_bk_off(0, poly, 0, j, part, 0));
            }
}
// one step: acc <- BK (x) src + acc with src = X^abar acc - acc (rotate) or acc <- BK (x) acc (plain external product)
static void t2_step(const uint32_t* dev, uint32_t* acc, bool rotate, uint32_t abar, uint32_t mask) {
    std::vector<uint32_t> dh(6 * T2_TILE_WORDS);
    for (int pw = 0; pw < 2; pw++) {
        uint32_t u[32][32];
        for (int lane = 0; lane < 32; lane++) {
            if (rotate) t2_u<true>(lane, acc + pw * 1024, abar, mask, u[lane]);
            else t2_u<false>(lane, acc + pw * 1024, abar, mask, u[lane]);
        }
        for (int dw = 0; dw < 3; dw++) {
            uint32_t* S = dh.data() + (3 * pw + dw) * T2_TILE_WORDS;
            for (int lane = 0; lane < 32; lane++) t2_fwd_cols(lane, u[lane], 6 * dw, S, h_digit_tab2.v);
            for (int lane = 0; lane < 32; lane++) t2_fwd_rows(lane, S, TwRow{h_fwdB + lane * TWB_STRIDE});
        }
    }
    std::vector<uint32_t> T(2 * 2 * T2_TILE_WORDS);
    for (int pw = 0; pw < 2; pw++) {
        for (int lane = 0; lane < 32; lane++) {
            uint32_t y0[32], y1[32];
            t2_mac(lane, dev + (size_t)pw * (T2_STEP_WORDS / 2), dh.data(), TwRow{h_invB + lane * TWB_STRIDE}, y0, y1);
            t2_inv_store(lane, y0, T.data() + (2 * pw) * T2_TILE_WORDS);
            t2_inv_store(lane, y1, T.data() + (2 * pw + 1) * T2_TILE_WORDS);
        }
    }
    for (int pw = 0; pw < 2; pw++)
        for (int lane = 0; lane < 32; lane++) {
            uint32_t z[32];
            for (int r = 0; r < 32; r++) z[r] = rotate ? acc[pw * 1024 + 32 * r + lane] : 0u;
            for (int s = 0; s < 2; s++) t2_inv_cols(lane, T.data() + (2 * pw + s) * T2_TILE_WORDS, 16 * s, z);
            for (int r = 0; r < 32; r++) acc[pw * 1024 + 32 * r + lane] = z[r];
        }
}
void emul_external_product_t2(const uint32_t* dev, const uint32_t* trlwe, uint32_t mask, uint32_t* out) {
    memcpy(out, trlwe, 2 * 1024 * sizeof(uint32_t));
    t2_step(dev, out, false, 0, mask);
}
void emul_cmux_rotate_t2(const uint32_t* dev, uint32_t* acc, uint32_t abar, uint32_t mask) { t2_step(dev, acc, true, abar, mask); }
// ---- FFT64 mode (fft64.cuh, blind_rotate_f64.cuh): one gate on ONE warp, f64 complex transform, exact rounding ----
// TRGSW torus polys [row j][poly o][1024] -> [row j][o][register][lane] complex, scaled by 1/512
static void f64_forward_all_lanes(cd (*x)[16], cd16* S, cd (*y)[16]) {
    const cd16* tb = reinterpret_cast<const cd16*>(h_f64_fwdB);
    static_assert(sizeof(cd16) == 16, "cd16 layout");
    for (int lane = 0; lane < 32; lane++) { f64_fwd_passA(x[lane]); f64_t1_store(lane, x[lane], S); }
    for (int lane = 0; lane < 32; lane++) {
        F64TwB tw;
        f64_fwd_twB(lane, tb, tw);
        f64_t1_load_cross(lane, S, tw.w[0], y[lane]);
        f64_fwd_passB(y[lane], tw);
    }
}
void emul_key_transform_f64(const uint32_t* trgsw, double* dev) {
    std::vector<cd16> S(512);
    cd16* out = reinterpret_cast<cd16*>(dev);
    for (int j = 0; j < BK_ROWS; j++)
        for (int o = 0; o < 2; o++) {
            cd x[32][16], y[32][16];
            for (int lane = 0; lane < 32; lane++) f64_key_input(lane, trgsw + (size_t)(j * 2 + o) * 1024, x[lane]);
            f64_forward_all_lanes(x, S.data(), y);
            for (int lane = 0; lane < 32; lane++)
                for (int k = 0; k < 16; k++) {
                    cd16 v; v.re = y[lane][k].re * F64_KEY_SCALE; v.im = y[lane][k].im * F64_KEY_SCALE;
                    out[f64_key_off(0, j, o) + k * 32 + lane] = v;
                }
        }
}
// one step: acc <- BK (x) src + acc with src = X^abar acc - acc (rotate), or acc <- BK (x) acc (plain external product).
// Returns the largest distance of a pre-rounding value from the nearest integer (the exactness margin: must stay far below 1/2).
static double f64_step(const double* dev, uint32_t* acc, bool rotate, uint32_t abar, uint32_t mask) {
    const cd16* key = reinterpret_cast<const cd16*>(dev);
    std::vector<cd16> S(512);
    std::vector<u4> planes(6 * 32);   // [digit][re / im][lane]
    cd sum[2][32][16];
    double frac = 0;
    for (int pw = 0; pw < 2; pw++) {
        for (int lane = 0; lane < 32; lane++) {
            uint32_t u[32];
            if (rotate) t2_u<true>(lane, acc + pw * 1024, abar, mask, u);
            else t2_u<false>(lane, acc + pw * 1024, abar, mask, u);
            f64_pack_plane<0>(u, planes[0 * 32 + lane], planes[1 * 32 + lane]);
            f64_pack_plane<1>(u, planes[2 * 32 + lane], planes[3 * 32 + lane]);
            f64_pack_plane<2>(u, planes[4 * 32 + lane], planes[5 * 32 + lane]);
        }
        for (int dw = 0; dw < 3; dw++) {
            cd x[32][16], y[32][16];
            for (int lane = 0; lane < 32; lane++) f64_digits(planes[(2 * dw) * 32 + lane], planes[(2 * dw + 1) * 32 + lane], x[lane]);
            f64_forward_all_lanes(x, S.data(), y);
            const int j = 3 * pw + dw;
            for (int o = 0; o < 2; o++)
                for (int lane = 0; lane < 32; lane++) {
                    if (j == 0) f64_mul(lane, y[lane], key + f64_key_off(0, j, o), sum[o][lane]);
                    else f64_mac(lane, y[lane], key + f64_key_off(0, j, o), sum[o][lane]);
                }
        }
    }
    const cd16* ta = reinterpret_cast<const cd16*>(h_f64_invA);
    const cd16* ut = reinterpret_cast<const cd16*>(h_f64_untw);
    for (int o = 0; o < 2; o++) {
        cd send[32][8], v[32][16], w[32][16];
        for (int lane = 0; lane < 32; lane++) { f64_inv_low(sum[o][lane]); f64_x_send(lane, sum[o][lane], send[lane]); }
        for (int lane = 0; lane < 32; lane++) { f64_inv_x_bfly(lane, sum[o][lane], send[lane ^ 1], v[lane]); f64_t2_store(lane, v[lane], S.data()); }
        for (int lane = 0; lane < 32; lane++) {
            uint32_t lo[16], hi[16];
            f64_t2_load(lane, S.data(), w[lane]);
            f64_inv_passA(lane, w[lane], ta);
            f64_untwist_round(lane, w[lane], ut, lo, hi, &frac);
            for (int r = 0; r < 16; r++) {
                uint32_t* a0 = acc + o * 1024 + 32 * r + lane;
                a0[0] = (rotate ? a0[0] : 0u) + lo[r];
                a0[512] = (rotate ? a0[512] : 0u) + hi[r];
            }
        }
    }
    return frac;
}
double emul_external_product_f64(const double* dev, const uint32_t* trlwe, uint32_t mask, uint32_t* out) {
    memcpy(out, trlwe, 2 * 1024 * sizeof(uint32_t));
    return f64_step(dev, out, false, 0, mask);
}
double emul_cmux_rotate_f64(const double* dev, uint32_t* acc, uint32_t abar, uint32_t mask) { return f64_step(dev, acc, true, abar, mask); }
// ---- FFT64 latency kernel (blind_rotate_f64l2.cuh): one transform on two warps (64 threads x 8 values), three radix-8 passes.
// Plain external product with the key in the one-warp layout (emul_key_transform_f64): the same transposes, swizzle and tables
// as the kernel, thread by thread.
static void l2_forward64(cd (*x)[8], cd16* A, cd16* Bb) {   // x[t][e] = z_{t + 64 e}  ->  x[t][e] = spectrum position 8 t + e
    const cd16* tf2 = reinterpret_cast<const cd16*>(h_l2_fwd2);
    const cd16* tf3 = reinterpret_cast<const cd16*>(h_l2_fwd3);
    for (int t = 0; t < 64; t++) {
        l2_fwd_pass1(x[t]);
        for (int e = 0; e < 8; e++) { A[t + 64 * e].re = x[t][e].re; A[t + 64 * e].im = x[t][e].im; }
    }
    for (int t = 0; t < 64; t++) {
        const int hi3 = t >> 3, lo3 = t & 7;
        cd16 w[4];
        for (int k = 0; k < 4; k++) w[k] = tf2[k * 8 + hi3];
        for (int m = 0; m < 8; m++) { x[t][m].re = A[64 * hi3 + 8 * m + lo3].re; x[t][m].im = A[64 * hi3 + 8 * m + lo3].im; }
        l2_fwd_pass23(x[t], w);
        for (int m = 0; m < 8; m++) { Bb[64 * hi3 + 8 * m + (lo3 ^ m)].re = x[t][m].re; Bb[64 * hi3 + 8 * m + (lo3 ^ m)].im = x[t][m].im; }
    }
    for (int t = 0; t < 64; t++) {
        const int lo3 = t & 7;
        cd16 w[4];
        for (int k = 0; k < 4; k++) w[k] = tf3[k * 64 + t];
        for (int e = 0; e < 8; e++) { x[t][e].re = Bb[8 * t + (e ^ lo3)].re; x[t][e].im = Bb[8 * t + (e ^ lo3)].im; }
        l2_fwd_pass23(x[t], w);
    }
}
static void l2_inverse64(cd (*y)[8], cd16* A, cd16* Bb, uint32_t* out /*[1024]*/) {   // y[t][e] = spectrum position 8 t + e
    const cd16* ti2 = reinterpret_cast<const cd16*>(h_l2_inv2);
    const cd16* ti3 = reinterpret_cast<const cd16*>(h_l2_inv3);
    const cd16* tut = reinterpret_cast<const cd16*>(h_l2_untw);
    for (int t = 0; t < 64; t++) {
        const int lo3 = t & 7;
        l2_inv_pass1(y[t]);
        for (int e = 0; e < 8; e++) { Bb[8 * t + (e ^ lo3)].re = y[t][e].re; Bb[8 * t + (e ^ lo3)].im = y[t][e].im; }
    }
    for (int t = 0; t < 64; t++) {
        const int hi3 = t >> 3, lo3 = t & 7;
        cd16 v[4];
        for (int k = 0; k < 4; k++) v[k] = ti2[k * 8 + lo3];
        for (int m = 0; m < 8; m++) { y[t][m].re = Bb[64 * hi3 + 8 * m + (lo3 ^ m)].re; y[t][m].im = Bb[64 * hi3 + 8 * m + (lo3 ^ m)].im; }
        l2_inv_pass23(y[t], v);
        for (int m = 0; m < 8; m++) { A[64 * hi3 + 8 * m + lo3].re = y[t][m].re; A[64 * hi3 + 8 * m + lo3].im = y[t][m].im; }
    }
    for (int t = 0; t < 64; t++) {
        cd16 v[4];
        for (int k = 0; k < 4; k++) v[k] = ti3[k * 64 + t];
        for (int e = 0; e < 8; e++) { y[t][e].re = A[t + 64 * e].re; y[t][e].im = A[t + 64 * e].im; }
        l2_inv_pass23(y[t], v);
        for (int e = 0; e < 8; e++) {
            const cd16 u = tut[e * 64 + t];
            const double zr = F_FMA(y[t][e].re, u.re, -F_MUL(y[t][e].im, u.im));
            const double zi = F_FMA(y[t][e].re, u.im, F_MUL(y[t][e].im, u.re));
            out[t + 64 * e] = f64_low_word(F_ADD(zr, F64_ROUND_MAGIC));
            out[512 + t + 64 * e] = f64_low_word(F_ADD(zi, F64_ROUND_MAGIC));
        }
    }
}
void emul_external_product_f64l2(const double* dev, const uint32_t* trlwe, uint32_t mask, uint32_t* out) {
    const cd16* key = reinterpret_cast<const cd16*>(dev);
    std::vector<cd16> A(512), Bb(512);
    std::vector<cd16> prod(12 * 512);   // [row j][output o][register e][thread t]
    for (int pair = 0; pair < 6; pair++) {
        const int pw = pair / 3, dw = pair % 3, sh = 6 * dw;
        cd x[64][8];
        for (int t = 0; t < 64; t++)
            for (int e = 0; e < 8; e++) {
                const uint32_t ur = add_alu(trlwe[pw * 1024 + t + 64 * e], mask) ^ mask;
                const uint32_t ui = add_alu(trlwe[pw * 1024 + 512 + t + 64 * e], mask) ^ mask;
                x[t][e].re = (double)((((int32_t)(ur << sh)) >> 24) & ~3);
                x[t][e].im = (double)((((int32_t)(ui << sh)) >> 24) & ~3);
            }
        l2_forward64(x, A.data(), Bb.data());
        for (int t = 0; t < 64; t++)
            for (int e = 0; e < 8; e++)
                for (int o = 0; o < 2; o++) {
                    const int pp = 8 * t + e;   // the kernel reads the same value from the [register 8][thread 64] copy of the key
                    const cd16 k = key[f64_key_off(0, pair, o) + (size_t)(pp & 15) * 32 + (pp >> 4)];
                    cd16 q;
                    q.re = F_FMA(x[t][e].re, k.re, -F_MUL(x[t][e].im, k.im));
                    q.im = F_FMA(x[t][e].re, k.im, F_MUL(x[t][e].im, k.re));
                    prod[((size_t)(2 * pair + o) * 8 + e) * 64 + t] = q;
                }
    }
    for (int o = 0; o < 2; o++) {
        cd y[64][8];
        for (int t = 0; t < 64; t++)
            for (int e = 0; e < 8; e++) {
                cd16 v = prod[((size_t)o * 8 + e) * 64 + t];
                for (int j = 1; j < 6; j++) {
                    const cd16 q = prod[((size_t)(2 * j + o) * 8 + e) * 64 + t];
                    v.re = F_ADD(v.re, q.re); v.im = F_ADD(v.im, q.im);
                }
                y[t][e].re = v.re; y[t][e].im = v.im;
            }
        l2_inverse64(y, A.data(), Bb.data(), out + o * 1024);
    }
}
// one 64-bit word of the ChaCha20 block the production generator is built on (RFC 8439 known-answer test)
uint64_t emul_chacha20_u64(const uint32_t* key, uint64_t counter, uint64_t nonce, int lane8) { return tfhe_rng::chacha20_u64(key, counter, nonce, lane8); }
uint32_t emul_prime(void) { return P; }
int32_t emul_key_slice(uint32_t c, int part) { return key_slice(c, part); }
int32_t emul_key_slice2(uint32_t c, int part) { return key_slice(c, part, 2); }
}
