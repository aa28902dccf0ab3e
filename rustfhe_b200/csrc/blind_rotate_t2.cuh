// blind_rotate_t2.cuh -- K5T, the THROUGHPUT shape of the blind rotation (included by engine.cu only).
//   gate pre-combination + 635 x CMUX + sample extract (tfhe.rs:27-113, trgsw.rs:264-322, trlwe.rs:110-121)
// One gate = TWO warps (warp pw owns accumulator polynomial pw), two 16-bit key slices, G gates per CTA (default 6: 12 warps =
// 3 per SM sub-partition, up to 168 registers per thread, 33.3 KB of shared memory per gate).  Per step and gate: two
// 64-thread named barriers and nothing else; the gates of a CTA never synchronise with each other.  See t2_steps.cuh for the
// work split and DESIGN.md section 3 for the measurements that led here.
#pragma once
#include "blind_rotate.cuh"
#include "t2_steps.cuh"

constexpr int T2_WARPS_PER_GATE = 2;
constexpr int T2_THREADS_PER_GATE = 64;
constexpr int T2_GATE_SMEM_WORDS = 2 * 1024 /*acc*/ + 6 * T2_TILE_WORDS /*digit spectra / transpose scratch*/ + 320 /*abar u16[640]*/;
constexpr size_t t2_smem_bytes(int G) { return (size_t)(TW_SMEM_WORDS + G * T2_GATE_SMEM_WORDS) * 4; }
static_assert(t2_smem_bytes(6) + 1024 <= 227 * 1024, "six gates must fit the shared memory of one SM");

// row twiddles of one lane held in registers for all transforms of a step (62 words: 16 LDS.128 once instead of per transform)
struct TwRegs {
    uint32_t v[64];
    __device__ __forceinline__ void load(const uint32_t* row) {
#pragma unroll
        for (int q = 0; q < 16; q++) {
            const uint4 t = *reinterpret_cast<const uint4*>(row + 4 * q);
            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
        }
    }
    __device__ __forceinline__ void get(int q, uint32_t& w, uint32_t& ws) const { w = v[2 * q]; ws = v[2 * q + 1]; }
    __device__ __forceinline__ void get2(int q, uint32_t& w0, uint32_t& ws0, uint32_t& w1, uint32_t& ws1) const {
        w0 = v[2 * q]; ws0 = v[2 * q + 1]; w1 = v[2 * q + 2]; ws1 = v[2 * q + 3];
    }
};

// TWREG bit 0: forward row twiddles in registers for the three forward transforms of a step; bit 1: same for the inverse.
// EXTPROD: one plain external product per "gate" (trgsw.rs:264-306) or cmux (trgsw.rs:315-322): out = BK[g % ntrgsw] (x) (in - in0) + in0.
template <int G, int TWREG, bool EXTPROD = false>
__global__ void __launch_bounds__(G* T2_THREADS_PER_GATE, 1) blind_rotate_t2_kernel(const BrArgs a) {
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* twF = smem;
    uint32_t* twI = smem + 32 * TWB_STRIDE;
    uint32_t* dtab = smem + 2 * 32 * TWB_STRIDE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gl = warp >> 1, pw = warp & 1;
    const int tid2 = threadIdx.x - gl * T2_THREADS_PER_GATE;
    uint32_t* acc = smem + TW_SMEM_WORDS + gl * T2_GATE_SMEM_WORDS;
    uint32_t* dh = acc + 2 * 1024;
    uint16_t* abar = reinterpret_cast<uint16_t*>(dh + 6 * T2_TILE_WORDS);

    // gates are dealt out evenly: the first cta_rem CTAs own cta_base+1 consecutive gates, the others cta_base (<= G)
    const long cta = blockIdx.x;
    const long first = cta * a.cta_base + (cta < a.cta_rem ? cta : a.cta_rem);
    const int cnt = a.cta_base + (cta < a.cta_rem ? 1 : 0);
    const bool active = gl < cnt;
    const long gate = active ? first + gl : a.B - 1;

    for (int t = threadIdx.x; t < 32 * TWB_STRIDE; t += blockDim.x) {
        twF[t] = g_fwdB[t];
        twI[t] = g_invB[t];
    }
    for (int t = threadIdx.x; t < DIGIT_TAB_WORDS; t += blockDim.x) dtab[t] = g_digit_tab2.v[t];
    // ---- prologue: gate pre-combination (tfhe.rs:27-71), rounding of (b, a) (tfhe.rs:97,107-108), acc_0 ----
    if (EXTPROD) {
        const uint32_t* src = a.trlwe_in + (size_t)gate * 2048;
        const uint32_t* sub = a.trlwe_in0 ? a.trlwe_in0 + (size_t)gate * 2048 : nullptr;   // cmux: rep_1 - rep_0
        for (int k = tid2; k < 2048; k += T2_THREADS_PER_GATE) acc[k] = sub ? src[k] - sub[k] : src[k];
    } else {
        uint32_t* lin = dh;
        const bool second = gate >= a.split;
        const long gsrc = second ? gate - a.split : gate;
        const uint32_t* q0 = second ? a.in0b : a.in0;
        const uint32_t* q1 = second ? a.in1b : a.in1;
        uint32_t k0 = (uint32_t)(second ? a.c0b : a.c0), k1 = (uint32_t)(second ? a.c1b : a.c1), kb = second ? a.cbb : a.cb;
        if (a.ops) gate_coeffs(a.ops[gate], a.mu, k0, k1, kb);
        const uint32_t* p0 = q0 + (size_t)(a.idx0 ? (long)a.idx0[gate] : gsrc) * (LWE_N + 1);
        const uint32_t* p1 = (q1 && k1 != 0) ? q1 + (size_t)(a.idx1 ? (long)a.idx1[gate] : gsrc) * (LWE_N + 1) : nullptr;
        for (int c = tid2; c <= LWE_N; c += T2_THREADS_PER_GATE) {
            uint32_t v = k0 * p0[c];
            if (p1) v += k1 * p1[c];
            if (c == 0) v += kb;
            lin[c] = v;
        }
        __syncthreads();
        for (int i = tid2; i < LWE_N; i += T2_THREADS_PER_GATE) abar[i] = (uint16_t)((lin[1 + i] + (1u << 20)) >> 21);  // round
        const uint32_t bbar = lin[0] >> 21;                                                                           // floor
        const uint32_t nrot = (2048u - bbar) & 2047u;  // acc_0 = X^{-bbar} * (mu, ..., mu ; 0)
        for (int k = tid2; k < 1024; k += T2_THREADS_PER_GATE) {
            const bool neg = ((uint32_t)k < (nrot & 1023u)) != (nrot >= 1024u);
            acc[k] = neg ? 0u - a.mu : a.mu;
            acc[1024 + k] = 0;
        }
    }
    __syncthreads();
    if (!active) return;   // gate slots without a gate leave here: every barrier below is private to one gate
    const int bar_gate = 1 + gl;
    uint32_t* accp = acc + pw * 1024;
    uint32_t* own = dh + 3 * pw * T2_TILE_WORDS;
    const uint32_t* keyp = a.bkdev + (size_t)pw * (T2_STEP_WORDS / 2) + (EXTPROD ? (size_t)(gate % a.ntrgsw) * T2_STEP_WORDS : (size_t)0);
    const int nsteps = EXTPROD ? 1 : a.nsteps;
    // ---- 635 x CMUX ----
#pragma unroll 1
    for (int i = 0; i < nsteps; i++, keyp += T2_STEP_WORDS) {
        {   // phase 1: u in registers, three forward transforms into the own spectrum tiles
            uint32_t u[32];
            t2_u<!EXTPROD>(lane, accp, EXTPROD ? 0u : (uint32_t)abar[i], a.mask, u);
            if (TWREG & 1) {
                TwRegs tw;
                tw.load(twF + lane * TWB_STRIDE);
#pragma unroll 1
                for (int dw = 0; dw < 3; dw++) {
                    uint32_t* S = own + dw * T2_TILE_WORDS;
                    t2_fwd_cols(lane, u, 6 * dw, S, dtab);
                    __syncwarp();
                    t2_fwd_rows(lane, S, tw);
                }
            } else {
#pragma unroll 1
                for (int dw = 0; dw < 3; dw++) {
                    uint32_t* S = own + dw * T2_TILE_WORDS;
                    t2_fwd_cols(lane, u, 6 * dw, S, dtab);
                    __syncwarp();
                    t2_fwd_rows(lane, S, TwRow{twF + lane * TWB_STRIDE});
                }
            }
        }
#if !defined(T2_EXP_NOBAR)
        bar_sync(bar_gate, T2_THREADS_PER_GATE);   // the six spectra of this step are complete
#endif
        {
            uint32_t y0[32], y1[32];
            if (TWREG & 2) {
                TwRegs tw;
                tw.load(twI + lane * TWB_STRIDE);
                t2_mac(lane, keyp, dh, tw, y0, y1);
            } else {
                t2_mac(lane, keyp, dh, TwRow{twI + lane * TWB_STRIDE}, y0, y1);
            }
#if !defined(T2_EXP_NOBAR)
            bar_sync(bar_gate, T2_THREADS_PER_GATE);   // both warps have read the spectra: the own tiles are scratch now
#endif
            t2_inv_store(lane, y0, own);
            t2_inv_store(lane, y1, own + T2_TILE_WORDS);
        }
        __syncwarp();
        {   // phase 3: inverse column passes of both slices, acc[pw] += x0 + (x1 << 16)
            uint32_t z[32];
#pragma unroll
            for (int r = 0; r < 32; r++) z[r] = EXTPROD ? 0u : accp[32 * r + lane];   // the plain product REPLACES the accumulator
#pragma unroll 1
            for (int s = 0; s < 2; s++) t2_inv_cols(lane, own + s * T2_TILE_WORDS, 16 * s, z);
#pragma unroll
            for (int r = 0; r < 32; r++) accp[32 * r + lane] = z[r];
        }
        __syncwarp();   // acc[pw] is complete before the next step's rotated reads (other lanes' rows)
    }
    bar_sync(bar_gate, T2_THREADS_PER_GATE);

    // ---- epilogue: sample_extract_index(0) (trlwe.rs:110-121) + key-switch digits (tlwe.rs:47-64) ----
    if (a.trlwe_out) {
        uint32_t* dst = a.trlwe_out + (size_t)gate * 2048;
        const uint32_t* add = (EXTPROD && a.trlwe_in0) ? a.trlwe_in0 + (size_t)gate * 2048 : nullptr;   // cmux: ... + rep_0
        for (int k = tid2; k < 2048; k += T2_THREADS_PER_GATE) dst[k] = add ? acc[k] + add[k] : acc[k];
    }
    if (a.ksdig || a.lwe1_out) {
        for (int i = tid2; i < 1024; i += T2_THREADS_PER_GATE) {
            const uint32_t ai = (i == 0) ? acc[1024] : 0u - acc[1024 + 1024 - i];
            if (a.ksdig) a.ksdig[(size_t)gate * 1024 + i] = (uint16_t)((ai + 0x8000u) >> 16);
            if (a.lwe1_out) a.lwe1_out[(size_t)gate * 1025 + 1 + i] = ai;
        }
        if (a.lwe1_out && tid2 == 0) a.lwe1_out[(size_t)gate * 1025] = acc[0];
    }
    if (a.out_init) {
        uint32_t* dst = a.out_init + (size_t)(a.idxo ? (long)a.idxo[gate] : gate) * (LWE_N + 1);
        for (int c = tid2; c <= LWE_N; c += T2_THREADS_PER_GATE) dst[c] = (c == 0) ? acc[0] : 0u;
    }
}

// K8T: key transform into the throughput layout.  One warp per (step i, row j, poly); loops over the two slices.
__global__ void __launch_bounds__(KT_WARPS * 32) bk_transform_t2_kernel(const uint32_t* __restrict__ bk, uint32_t* __restrict__ dev,
                                                                       int npolys /* = nsteps*12 */) {
    __shared__ __align__(16) uint32_t twF[32 * TWB_STRIDE];
    __shared__ __align__(16) uint32_t scratch[KT_WARPS][TILE_WORDS];
    for (int t = threadIdx.x; t < 32 * TWB_STRIDE; t += blockDim.x) twF[t] = g_fwdB[t];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pid = blockIdx.x * KT_WARPS + warp;
    if (pid >= npolys) return;
    const int poly = pid & 1, j = (pid >> 1) % BK_ROWS, i = pid / (2 * BK_ROWS);
    const uint32_t* src = bk + (size_t)pid * 1024;
    uint32_t* S = scratch[warp];
    for (int part = 0; part < 2; part++) {
        key_cols(lane, src, part, S, 2);
        __syncwarp();
        key_rows_t2(lane, S, twF, dev + t2_bk_off(i, poly, 0, j, part, 0));
        __syncwarp();
    }
}
