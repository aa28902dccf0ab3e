// engine.cu -- host side and C ABI (include/tfhe_b200.h) of the B200 TFHE gate-bootstrapping engine; the one translation unit
// that is compiled with nvcc for sm_100a.  Device code (DESIGN.md has the roofline of each kernel):
//   blind_rotate.cuh : K8 bk_transform_kernel, K5 blind_rotate_kernel, K5L blind_rotate_pair_kernel
//   keyswitch.cuh    : K6 keyswitch2_kernel (default), keyswitch_kernel, lwe1_prepare_kernel
//   aux_kernels.cuh  : polymul_kernel, device keygen / encryption / decryption / sample-extract kernels
//   cmux_steps.cuh, ntt32.cuh, tfhe_rng.cuh : per-lane arithmetic shared with the CPU emulation and the host keygen
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <climits>
#include <cstdint>
#include <algorithm>
#include <string>
#include <vector>
#include "../../include/tfhe_b200.h"
#include "blind_rotate.cuh"
#include "blind_rotate_t2.cuh"
#include "blind_rotate_f64.cuh"
#include "blind_rotate_f64t.cuh"
#include "blind_rotate_f64l2.cuh"
#include "blind_rotate_f64w2.cuh"
#include "keyswitch.cuh"
#include "aux_kernels.cuh"

using namespace tfhe;

// =====================================================================================================
// host side: context + C ABI
// =====================================================================================================
// Work slots: every gate-batch call borrows one slot (key-switch digit workspace, staging buffers, scratch) from a
// small ring.  A slot is guarded by an event recorded behind the last operation that touches its buffers, so calls
// issued on DIFFERENT streams (or back-to-back asynchronous host calls) may overlap on the device: the tail of one
// batch's blind rotation and its key switch run under the head of the next batch.
struct Slot {
    cudaStream_t stream = nullptr;   // internal stream, used by the host-pointer entry points
    cudaEvent_t done = nullptr;
    bool pending = false;
    cudaStream_t last = nullptr;     // the stream `done` was last recorded on
    uint16_t* ksdig = nullptr; size_t ksdig_cap = 0;
    uint32_t* tmp[4] = {nullptr, nullptr, nullptr, nullptr}; size_t tmp_cap[4] = {0, 0, 0, 0};
    uint32_t* scratch = nullptr; size_t scratch_cap = 0;   // hom_mux intermediates / transformed TRGSWs of step-level calls
    uint8_t* s0buf = nullptr;                               // [1024] device copy of a caller's lv0 secret key (encrypt / decrypt)
    uint8_t* opsbuf = nullptr; size_t opsbuf_cap = 0;       // per-gate opcodes of a mixed batch (host-pointer entry)
};

static const size_t BK_TORUS_BYTES = (size_t)LWE_N * 12 * 1024 * 4;
static const size_t KSK_BYTES = (size_t)1024 * 8 * 3 * (LWE_N + 1) * 4;
struct tfhe_b200_ctx {
    tfhe_b200_params prm;
    int device = 0;
    int sm_count = 0;
    uint32_t* bkdev = nullptr;   // n * BK_STEP_WORDS
    uint32_t* bkdev_t2 = nullptr;   // n * T2_STEP_WORDS: the two-slice key in the throughput kernel's layout (blind_rotate_t2.cuh)
    cd16* bkdev_f64 = nullptr;      // n * F64_STEP_ELEMS: the f64 spectra of the FFT64 mode (blind_rotate_f64.cuh), allocated on first use
    cd16* bkdev_f64l = nullptr;     // the same values in the layout of the latency kernel (blind_rotate_f64l2.cuh)
    uint32_t* kskdev = nullptr;  // [N][t][3][n+1]
    uint32_t* bk_torus = nullptr;  // [n][2l][2][N] torus-domain key as loaded / generated (kept for export: 31 MB)
    uint8_t* keybits = nullptr;    // device copy of (s0[n] | pad to 1024 | s1[N]) during device keygen
    int32_t* s1poly = nullptr;     // s1 as a polynomial of small integers, the multiplier of the a*s products
    bool have_bk = false, have_ksk = false;
    static constexpr int NSLOT = 4;
    Slot slots[NSLOT];
    unsigned next_slot = 0;
    cudaEvent_t keys_ev = nullptr;           // recorded behind the last asynchronous key load (load_*_device on a caller's stream);
    bool keys_pending = false;               // every later batch waits for it, whatever stream it is issued on
    static constexpr int RING = 64;          // event ring: per-launch device times of the last RING timed gate batches
    cudaEvent_t ev[RING][4] = {};
    uint64_t timed = 0;
    uint64_t launches = 0;
    uint64_t last_batch = 0;
    int gates_per_cta = 1;
    int variant = 7;  // blind-rotate launch shape, see launch_blind_rotate
    int key_slices = 1;  // arithmetic mode (tfhe_b200_set_key_slices): 1 (default) = FFT64, one f64 complex transform with exact rounding
                         // for every gate batch (the step-level entry points run the two-slice NTT); 2 = NTT, two
                         // 16-bit key slices; both exact for honestly generated keys (DESIGN.md section 2 has the margins);
                         // 3 = NTT, three 11-bit slices, exact in the worst case
    int ns_int() const { return key_slices == 3 ? 3 : 2; }   // slices of the integer (NTT) form of the key
    int f64_tmem = 0;           // FFT64 throughput kernel: 0 = K5F, eight gates per SM (default); 2 = K5F2, a gate on two warps, six gates; 1 = per-gate state in tensor memory, twelve gates
                                // per SM (K5FT, TFHE_B200_F64_TMEM=1): measured 10 % slower -- the kernel is bound by issue slots, not by latency
    int f64_cluster = 1;        // FFT64 latency shape on a cluster of two SMs for batches of at most #SMs/2 gates (TFHE_B200_F64_CLUSTER=0: one SM)
    int f64_latency = 1;        // FFT64 mode: batches of at most 3 #SMs gates run one gate per SM, two warps per transform (K5FL2);
                                // TFHE_B200_F64_LATENCY=0: the NTT latency shapes / K5T / K5F instead
    int f64_stagger_ns = 0;   // start-up offset between the warps of a CTA of the FFT64 kernel (TFHE_B200_F64_STAGGER)
    int t2_gates = 6;    // gates per CTA of the throughput kernel (TFHE_B200_T2_G: 4 or 6)
    int t2_twreg = 1;    // which row twiddles the throughput kernel keeps in registers (TFHE_B200_T2_TWREG: bit 0 forward, bit 1 inverse)
    int slab_tma = 1;     // one gate per CTA: key slabs staged by bulk copies (TFHE_B200_SLAB_TMA=0: streamed from L2 by the warps)
    int pair_max = 0;     // largest batch that runs on 2-SM clusters (set at create: #SMs / 2)
    int deal_fixed = -1;  // how a full batch is cut into CTAs: 0 = dealt evenly over whole waves (best for a batch running alone),
                          // 1 = 4-gate CTAs only (best when batches on other streams back-fill the last wave), -1 = decide per call
    int ks_variant = 3;  // key-switch kernel: 3 = rows staged in shared memory, one warp per gate, producer / consumer ring; 2 = the same with a CTA barrier per stage; 1 = register tiles
    std::string err;
};
static thread_local std::string g_create_err;

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) {                                                                          \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                \
            return TFHE_B200_ERR_CUDA;                                                                    \
        }                                                                                                 \
    } while (0)
#define RC(call)                                                                                          \
    do {                                                                                                  \
        int rc_ = (call);                                                                                 \
        if (rc_) return rc_;                                                                              \
    } while (0)

static int fail(tfhe_b200_ctx* ctx, int code, const char* msg) {
    if (ctx) ctx->err = msg; else g_create_err = msg;
    return code;
}
static int grow(tfhe_b200_ctx* ctx, void** p, size_t* cap, size_t bytes) {
    if (*cap >= bytes) return TFHE_B200_OK;
    if (*p) CK(cudaFree(*p));   // cudaFree waits for the device: nothing in flight can still use the old buffer
    *p = nullptr; *cap = 0;
    CK(cudaMalloc(p, bytes));
    *cap = bytes;
    return TFHE_B200_OK;
}
// borrow the next slot of the ring for work that will be enqueued on `st` (nullptr = the slot's own stream)
static int slot_acquire(tfhe_b200_ctx* ctx, cudaStream_t* st, bool own_stream, Slot** out) {
    Slot& s = ctx->slots[ctx->next_slot++ % tfhe_b200_ctx::NSLOT];
    if (own_stream) *st = s.stream;
    if (s.pending) CK(cudaStreamWaitEvent(*st, s.done, 0));
    if (ctx->keys_pending) {
        if (cudaEventQuery(ctx->keys_ev) == cudaSuccess) ctx->keys_pending = false;
        else CK(cudaStreamWaitEvent(*st, ctx->keys_ev, 0));
    }
    *out = &s;
    return TFHE_B200_OK;
}
static int keys_loaded_on(tfhe_b200_ctx* ctx, cudaStream_t st) {
    if (ctx->keys_pending) CK(cudaStreamWaitEvent(st, ctx->keys_ev, 0));   // chain: the new record also covers an earlier load
    CK(cudaEventRecord(ctx->keys_ev, st));
    ctx->keys_pending = true;
    return TFHE_B200_OK;
}
static int slot_release(tfhe_b200_ctx* ctx, Slot* s, cudaStream_t st) {
    CK(cudaEventRecord(s->done, st));
    s->pending = true;
    s->last = st;
    return TFHE_B200_OK;
}

extern "C" {
static int transform_keys(tfhe_b200_ctx* ctx, const uint32_t* src_dev, uint32_t* dst_dev, int nsteps, cudaStream_t st);
}
template <class Kern>
static cudaError_t set_smem(Kern k, int G) { return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)br_smem_bytes(G)); }

extern "C" {

const char* tfhe_b200_version(void) { return "rustfhe_b200 0.4 (sm_100a; FFT64 with exact rounding; NTT p=536856577 with 2x16-bit / 3x11-bit key slices selectable)"; }

int tfhe_b200_default_params(tfhe_b200_params* p) {
    if (!p) return TFHE_B200_ERR_PARAM;
    p->n = 635; p->N = 1024; p->l = 3; p->bgbit = 6; p->ks_t = 8; p->ks_basebit = 2;
    p->mu = 0x20000000u; p->decomp_mask = TFHE_B200_MASK_FAITHFUL;
    return TFHE_B200_OK;
}

int tfhe_b200_ctx_create(const tfhe_b200_params* p, int device, tfhe_b200_ctx** out) {
    if (!out) return fail(nullptr, TFHE_B200_ERR_PARAM, "ctx_create: out is NULL");
    *out = nullptr;
    tfhe_b200_params prm;
    tfhe_b200_default_params(&prm);
    if (p) prm = *p;
    if (prm.n != 635 || prm.N != 1024 || prm.l != 3 || prm.bgbit != 6 || prm.ks_t != 8 || prm.ks_basebit != 2)
        return fail(nullptr, TFHE_B200_ERR_PARAM, "ctx_create: this build supports n=635 N=1024 l=3 Bgbit=6 t=8 basebit=2 only");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_err = std::string("ctx_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback";
        return TFHE_B200_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) return fail(nullptr, TFHE_B200_ERR_PARAM, "ctx_create: bad device index");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_err = cudaGetErrorString(e); return TFHE_B200_ERR_CUDA; }
    if (prop.major != 10) {
        g_create_err = "ctx_create: device is not sm_100 (kernels are built for sm_100a only; no fallback)";
        return TFHE_B200_ERR_CUDA;
    }
    tfhe_b200_ctx* ctx = new tfhe_b200_ctx();
    ctx->prm = prm; ctx->device = device; ctx->sm_count = prop.multiProcessorCount;
    auto bail = [&](const char* what, cudaError_t ee) { g_create_err = std::string(what) + ": " + cudaGetErrorString(ee); tfhe_b200_ctx_destroy(ctx); return TFHE_B200_ERR_CUDA; };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
    for (auto& s : ctx->slots) {
        if ((e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
        if ((e = cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    }
    for (auto& slot : ctx->ev) for (auto& ev : slot) if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&ctx->keys_ev, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMalloc(&ctx->bkdev, (size_t)LWE_N * BK_STEP_WORDS * 4)) != cudaSuccess) return bail("cudaMalloc(bk)", e);
    if ((e = cudaMalloc(&ctx->bkdev_t2, (size_t)LWE_N * T2_STEP_WORDS * 4)) != cudaSuccess) return bail("cudaMalloc(bk t2)", e);
    if ((e = cudaMalloc(&ctx->kskdev, (size_t)1024 * 8 * 3 * (LWE_N + 1) * 4)) != cudaSuccess) return bail("cudaMalloc(ksk)", e);
    if ((e = cudaMalloc(&ctx->bk_torus, BK_TORUS_BYTES)) != cudaSuccess) return bail("cudaMalloc(bk_torus)", e);
    if ((e = cudaMalloc(&ctx->keybits, 2048)) != cudaSuccess) return bail("cudaMalloc(keybits)", e);
    if ((e = cudaMalloc(&ctx->s1poly, 1024 * 4)) != cudaSuccess) return bail("cudaMalloc(s1poly)", e);
    for (auto& s : ctx->slots) if ((e = cudaMalloc(&s.s0buf, 1024)) != cudaSuccess) return bail("cudaMalloc(s0buf)", e);
    if ((e = set_smem(blind_rotate_kernel<1, false, 1>, 1)) != cudaSuccess) return bail("smem attr", e);
    if ((e = set_smem(blind_rotate_kernel<1, false, 3>, 1)) != cudaSuccess) return bail("smem attr", e);
    if ((e = set_smem(blind_rotate_kernel<4, false, 1>, 4)) != cudaSuccess) return bail("smem attr", e);
    if ((e = set_smem(blind_rotate_kernel<2, true, 1>, 2)) != cudaSuccess) return bail("smem attr", e);
    if ((e = set_smem(blind_rotate_kernel<4, true, 1>, 4)) != cudaSuccess) return bail("smem attr", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_pair_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_WORDS * 4)) != cudaSuccess)
        return bail("smem attr (pair)", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM_WORDS * 4)) != cudaSuccess)
        return bail("smem attr (pair)", e);
    if ((e = set_smem(blind_rotate_kernel<4, false, 1, 2>, 4)) != cudaSuccess) return bail("smem attr", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_kernel<1, false, 1, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)br_smem_bytes_tma())) != cudaSuccess)
        return bail("smem attr (slab tma)", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_kernel<1, false, 1, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)br_smem_bytes_tma())) != cudaSuccess)
        return bail("smem attr (slab tma)", e);
    if ((e = set_smem(blind_rotate_kernel<1, false, 1, 2>, 1)) != cudaSuccess) return bail("smem attr", e);
    if ((e = set_smem(blind_rotate_kernel<2, true, 1, 2>, 2)) != cudaSuccess) return bail("smem attr", e);
    if ((e = set_smem(blind_rotate_kernel<4, true, 1, 2>, 4)) != cudaSuccess) return bail("smem attr", e);
    {
        auto t2attr = [&](auto kern, int G) { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)t2_smem_bytes(G)); };
        if ((e = t2attr(blind_rotate_t2_kernel<6, 0>, 6)) != cudaSuccess || (e = t2attr(blind_rotate_t2_kernel<6, 1>, 6)) != cudaSuccess ||
            (e = t2attr(blind_rotate_t2_kernel<6, 3>, 6)) != cudaSuccess || (e = t2attr(blind_rotate_t2_kernel<4, 0>, 4)) != cudaSuccess ||
            (e = t2attr(blind_rotate_t2_kernel<4, 1>, 4)) != cudaSuccess || (e = t2attr(blind_rotate_t2_kernel<4, 3>, 4)) != cudaSuccess ||
            (e = t2attr(blind_rotate_t2_kernel<6, 1, true>, 6)) != cudaSuccess)
            return bail("smem attr (t2)", e);
    }
    if (const char* v = getenv("TFHE_B200_F64_STAGGER")) ctx->f64_stagger_ns = atoi(v);
    if (const char* v = getenv("TFHE_B200_T2_G")) ctx->t2_gates = (atoi(v) == 4) ? 4 : 6;
    if (const char* v = getenv("TFHE_B200_T2_TWREG")) ctx->t2_twreg = atoi(v) & 3;
    if (const char* v = getenv("TFHE_B200_BR_VARIANT")) ctx->variant = atoi(v);
    if (const char* v = getenv("TFHE_B200_KS_VARIANT")) ctx->ks_variant = atoi(v);
    if (const char* v = getenv("TFHE_B200_DEAL_FIXED")) ctx->deal_fixed = atoi(v);
    ctx->pair_max = ctx->sm_count / 2;
    {   // clusters of two that the device can keep resident at once (GPCs with an odd number of SMs leave one unpaired): beyond
        // that a batch of clusters would need a second wave
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * (unsigned)ctx->pair_max); cfg.blockDim = dim3(PAIR_LAUNCH_THREADS); cfg.dynamicSmemBytes = (size_t)PAIR_SMEM_WORDS * 4;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nclusters = 0;
        if (cudaOccupancyMaxActiveClusters(&nclusters, blind_rotate_pair_kernel<3>, &cfg) == cudaSuccess && nclusters > 0)
            ctx->pair_max = std::min(ctx->pair_max, nclusters);
        else
            cudaGetLastError();
    }
    if (const char* v = getenv("TFHE_B200_SLAB_TMA")) ctx->slab_tma = atoi(v);
    if (const char* v = getenv("TFHE_B200_PAIR_MAX")) ctx->pair_max = std::min(atoi(v), ctx->sm_count / 2);
    if (const char* v = getenv("TFHE_B200_KEY_SLICES")) {   // A/B and test runs: say so, the arithmetic mode is not a silent setting
        const int k = atoi(v);
        if (k >= 1 && k <= 3 && k != ctx->key_slices) {
            ctx->key_slices = k;
            fprintf(stderr, "rustfhe_b200: TFHE_B200_KEY_SLICES=%d selects arithmetic mode %d (%s) for this context\n", k, k,
                    k == 1 ? "FFT64" : k == 2 ? "NTT, two key slices" : "NTT, three key slices");
        }
    }
    if ((e = cudaFuncSetAttribute(blind_rotate_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f64_smem_bytes())) != cudaSuccess)
        return bail("smem attr (f64)", e);
    if ((e = cudaFuncSetAttribute(polymul_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PMF_SMEM_BYTES)) != cudaSuccess)
        return bail("smem attr (f64 polymul)", e);
    if ((e = cudaFuncSetAttribute(external_product_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f64_smem_bytes())) != cudaSuccess)
        return bail("smem attr (f64 external product)", e);
    if ((e = cudaFuncSetAttribute(external_product_item_f64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, XPI_SMEM_BYTES)) != cudaSuccess)
        return bail("smem attr (f64 external product, per-item TRGSW)", e);
    if (const char* v = getenv("TFHE_B200_F64_LATENCY")) ctx->f64_latency = atoi(v);
    if (const char* v = getenv("TFHE_B200_F64_CLUSTER")) ctx->f64_cluster = atoi(v);
    if ((e = cudaFuncSetAttribute(blind_rotate_f64_latency3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F64L3_SMEM_BYTES)) != cudaSuccess)
        return bail("smem attr (f64 latency 3)", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_f64_latency2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F64L2_SMEM_BYTES)) != cudaSuccess)
        return bail("smem attr (f64 latency 2)", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_f64t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f64t_smem_bytes())) != cudaSuccess)
        return bail("smem attr (f64 tmem)", e);
    if ((e = cudaFuncSetAttribute(blind_rotate_f64w2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f64w2_smem_bytes())) != cudaSuccess)
        return bail("smem attr (f64 two warps)", e);
    if (const char* v = getenv("TFHE_B200_F64_TMEM")) ctx->f64_tmem = atoi(v);
    if (const char* v = getenv("TFHE_B200_F64_KERNEL"))   // "k5f" (default), "tmem" (K5FT), "w2" (K5F2): the measured alternatives
        ctx->f64_tmem = !strcmp(v, "tmem") ? 1 : !strcmp(v, "w2") ? 2 : 0;
    if ((e = cudaFuncSetAttribute(keyswitch2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KS2_SMEM_BYTES)) != cudaSuccess)
        return bail("smem attr (keyswitch2)", e);
    if ((e = cudaFuncSetAttribute(keyswitch_p_kernel<KSP_GATES, KSP_RING>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)KsP<KSP_GATES, KSP_RING>::SMEM_BYTES)) != cudaSuccess)
        return bail("smem attr (keyswitch pipeline)", e);
    *out = ctx;
    return TFHE_B200_OK;
}

int tfhe_b200_ctx_destroy(tfhe_b200_ctx* ctx) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    if (ctx->keybits) cudaMemset(ctx->keybits, 0, 2048);
    if (ctx->s1poly) cudaMemset(ctx->s1poly, 0, 1024 * 4);
    for (auto& s : ctx->slots) if (s.s0buf) cudaMemset(s.s0buf, 0, 1024);
    cudaFree(ctx->bkdev); cudaFree(ctx->bkdev_t2); cudaFree(ctx->bkdev_f64); cudaFree(ctx->bkdev_f64l); cudaFree(ctx->kskdev); cudaFree(ctx->bk_torus); cudaFree(ctx->keybits); cudaFree(ctx->s1poly);
    for (auto& s : ctx->slots) {
        cudaFree(s.ksdig); cudaFree(s.scratch); cudaFree(s.s0buf); cudaFree(s.opsbuf);
        for (auto p : s.tmp) cudaFree(p);
        if (s.done) cudaEventDestroy(s.done);
        if (s.stream) cudaStreamDestroy(s.stream);
    }
    for (auto& slot : ctx->ev) for (auto ev : slot) if (ev) cudaEventDestroy(ev);
    if (ctx->keys_ev) cudaEventDestroy(ctx->keys_ev);
    delete ctx;
    return TFHE_B200_OK;
}

const char* tfhe_b200_last_error(const tfhe_b200_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int tfhe_b200_set_decomp_mask(tfhe_b200_ctx* ctx, uint32_t mask) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    ctx->prm.decomp_mask = mask;
    return TFHE_B200_OK;
}

// Pre-allocate every work slot for batches of up to max_batch gates (hom_mux needs twice the key-switch workspace), so
// that no call in a latency- or throughput-critical region ever reaches cudaMalloc (which blocks while the GPU is busy).
int tfhe_b200_reserve(tfhe_b200_ctx* ctx, size_t max_batch) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    if (max_batch == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    const size_t ct = max_batch * (size_t)(LWE_N + 1) * 4;
    for (auto& s : ctx->slots) {
        RC(grow(ctx, (void**)&s.ksdig, &s.ksdig_cap, 2 * max_batch * 1024 * sizeof(uint16_t)));
        for (int k = 0; k < 4; k++) RC(grow(ctx, (void**)&s.tmp[k], &s.tmp_cap[k], ct));
        RC(grow(ctx, (void**)&s.scratch, &s.scratch_cap, 2 * ct));
    }
    return TFHE_B200_OK;
}

// Arithmetic mode of the gate path: 1 = FFT64 (default), 2 / 3 = NTT with two 16-bit / three 11-bit key slices (include/tfhe_b200.h
// has the exactness statement of each).  Re-transforms the loaded bootstrapping key.
int tfhe_b200_set_key_slices(tfhe_b200_ctx* ctx, int slices) {
    if (!ctx || slices < 1 || slices > 3) return fail(ctx, TFHE_B200_ERR_PARAM, "set_key_slices: 1 (FFT64), 2 or 3");
    if (slices == ctx->key_slices) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    ctx->key_slices = slices;
    if (ctx->have_bk) {
        cudaStream_t st = ctx->slots[0].stream;
        RC(transform_keys(ctx, ctx->bk_torus, ctx->bkdev, LWE_N, st));
        CK(cudaStreamSynchronize(st));
    }
    return TFHE_B200_OK;
}

int tfhe_b200_set_batch_overlap(tfhe_b200_ctx* ctx, int mode) {
    if (!ctx || mode < TFHE_B200_OVERLAP_AUTO || mode > TFHE_B200_OVERLAP_STREAMED) return fail(ctx, TFHE_B200_ERR_PARAM, "set_batch_overlap: -1, 0 or 1");
    ctx->deal_fixed = mode;
    return TFHE_B200_OK;
}

int tfhe_b200_sync(tfhe_b200_ctx* ctx) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    CK(cudaSetDevice(ctx->device));
    for (auto& s : ctx->slots)
        if (s.pending) { CK(cudaEventSynchronize(s.done)); s.pending = false; }
    return TFHE_B200_OK;
}

int tfhe_b200_reset_stats(tfhe_b200_ctx* ctx) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    ctx->timed = 0;
    return TFHE_B200_OK;
}
int tfhe_b200_get_stats(tfhe_b200_ctx* ctx, tfhe_b200_stats* out) {
    if (!ctx || !out) return TFHE_B200_ERR_PARAM;
    memset(out, 0, sizeof *out);
    out->kernel_launches = ctx->launches;
    out->last_batch = ctx->last_batch;
    out->gates_per_cta = ctx->gates_per_cta;
    out->sm_count = ctx->sm_count;
    out->key_slices = ctx->key_slices;
    // what the gate path streams: the two-slice key in the throughput layout (8 B per coefficient) or the three-slice key (12 B)
    out->device_key_bytes = (uint64_t)LWE_N * (ctx->key_slices == 1 ? F64_STEP_ELEMS * 4 /* one of the two layouts is streamed per launch */ : ctx->key_slices == 2 ? T2_STEP_WORDS : bk_step_words(3)) * 4 +
                            (uint64_t)1024 * 8 * 3 * (LWE_N + 1) * 4;
    const uint64_t cnt = ctx->timed < (uint64_t)tfhe_b200_ctx::RING ? ctx->timed : (uint64_t)tfhe_b200_ctx::RING;
    double sb = 0, sk = 0;
    for (uint64_t k = 0; k < cnt; k++) {
        const int slot = (int)((ctx->timed - 1 - k) % tfhe_b200_ctx::RING);
        float b = 0, s = 0;
        CK(cudaEventSynchronize(ctx->ev[slot][3]));
        CK(cudaEventElapsedTime(&b, ctx->ev[slot][0], ctx->ev[slot][1]));
        CK(cudaEventElapsedTime(&s, ctx->ev[slot][2], ctx->ev[slot][3]));
        if (k == 0) { out->last_blind_rotate_ms = b; out->last_keyswitch_ms = s; }
        sb += b; sk += s;
    }
    out->timed_launches = cnt;
    if (cnt) { out->avg_blind_rotate_ms = (float)(sb / cnt); out->avg_keyswitch_ms = (float)(sk / cnt); }
    return TFHE_B200_OK;
}

// ---- keys ----
static int transform_keys(tfhe_b200_ctx* ctx, const uint32_t* src_dev, uint32_t* dst_dev, int nsteps, cudaStream_t st) {
    const int npolys = nsteps * 12;
    bk_transform_kernel<<<(npolys + KT_WARPS - 1) / KT_WARPS, KT_WARPS * 32, 0, st>>>(src_dev, dst_dev, npolys, ctx->ns_int());
    ctx->launches++;
    CK(cudaGetLastError());
    if (ctx->key_slices == 1 && dst_dev == ctx->bkdev) {   // the whole key: also as f64 spectra
        if (!ctx->bkdev_f64) CK(cudaMalloc(&ctx->bkdev_f64, (size_t)LWE_N * F64_STEP_ELEMS * sizeof(cd16)));
        bk_transform_f64_kernel<<<(npolys + KTF_WARPS - 1) / KTF_WARPS, KTF_WARPS * 32, 0, st>>>(src_dev, ctx->bkdev_f64, npolys);
        if (!ctx->bkdev_f64l) CK(cudaMalloc(&ctx->bkdev_f64l, (size_t)LWE_N * F64_STEP_ELEMS * sizeof(cd16)));
        f64l2_key_layout_kernel<<<(unsigned)npolys, 256, 0, st>>>(ctx->bkdev_f64, ctx->bkdev_f64l, (long)npolys);
        ctx->launches += 2;
        CK(cudaGetLastError());
    }
    if (ctx->ns_int() == 2 && dst_dev == ctx->bkdev) {   // the whole key: also in the throughput kernel's layout
        bk_transform_t2_kernel<<<(npolys + KT_WARPS - 1) / KT_WARPS, KT_WARPS * 32, 0, st>>>(src_dev, ctx->bkdev_t2, npolys);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return TFHE_B200_OK;
}
int tfhe_b200_load_bk_device(tfhe_b200_ctx* ctx, const uint32_t* bk_dev, void* stream) {
    if (!ctx || !bk_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "load_bk_device: null argument");
    CK(cudaSetDevice(ctx->device));
    if (bk_dev != ctx->bk_torus) CK(cudaMemcpyAsync(ctx->bk_torus, bk_dev, BK_TORUS_BYTES, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    RC(transform_keys(ctx, ctx->bk_torus, ctx->bkdev, LWE_N, (cudaStream_t)stream));
    ctx->have_bk = true;
    return keys_loaded_on(ctx, (cudaStream_t)stream);
}
int tfhe_b200_load_bk(tfhe_b200_ctx* ctx, const uint32_t* bk_host) {
    if (!ctx || !bk_host) return fail(ctx, TFHE_B200_ERR_PARAM, "load_bk: null argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    CK(cudaMemcpyAsync(ctx->bk_torus, bk_host, BK_TORUS_BYTES, cudaMemcpyHostToDevice, st));
    RC(tfhe_b200_load_bk_device(ctx, ctx->bk_torus, st));
    CK(cudaStreamSynchronize(st));
    return TFHE_B200_OK;
}
int tfhe_b200_load_ksk_device(tfhe_b200_ctx* ctx, const uint32_t* ksk_dev, void* stream) {
    if (!ctx || !ksk_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "load_ksk_device: null argument");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->kskdev, ksk_dev, (size_t)1024 * 8 * 3 * (LWE_N + 1) * 4, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    ctx->have_ksk = true;
    return keys_loaded_on(ctx, (cudaStream_t)stream);
}
int tfhe_b200_load_ksk(tfhe_b200_ctx* ctx, const uint32_t* ksk_host) {
    if (!ctx || !ksk_host) return fail(ctx, TFHE_B200_ERR_PARAM, "load_ksk: null argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    CK(cudaMemcpyAsync(ctx->kskdev, ksk_host, (size_t)1024 * 8 * 3 * (LWE_N + 1) * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    ctx->have_ksk = true;
    return TFHE_B200_OK;
}

}  // extern "C" (helpers below are internal)

// ---- launches ----
static void op_coeffs(int op, uint32_t mu, int32_t* c0, int32_t* c1, uint32_t* cb, bool* two) {
    *two = true;
    switch (op) {
    case TFHE_B200_NAND: *c0 = -1; *c1 = -1; *cb = mu; break;
    case TFHE_B200_AND: *c0 = 1; *c1 = 1; *cb = 0u - mu; break;
    case TFHE_B200_OR: *c0 = 1; *c1 = 1; *cb = mu; break;
    case TFHE_B200_XOR: *c0 = 2; *c1 = 2; *cb = 2u * mu; break;
    case TFHE_B200_NOT: *c0 = -1; *c1 = 0; *cb = 0; *two = false; break;
    case TFHE_B200_ANDNY: *c0 = -1; *c1 = 1; *cb = 0u - mu; break;
    default: *c0 = 1; *c1 = 0; *cb = 0; *two = false; break;
    }
}
// Is a batch issued earlier on ANOTHER stream still running?  Then the device is shared between batches and whatever a
// batch leaves idle in its last wave is taken by the next one: cut the batch into full 4-gate CTAs (measured at 1024 gates,
// two streams: 60.2 k gates/s against 57.3 k dealt evenly; alone it is the other way round, 52.3 k against 56.4 k).
static bool batches_overlap(tfhe_b200_ctx* ctx, cudaStream_t st) {
    if (ctx->deal_fixed >= 0) return ctx->deal_fixed != 0;
    for (Slot& s : ctx->slots)
        if (s.pending && s.last != st) {
            const cudaError_t q = cudaEventQuery(s.done);
            if (q == cudaErrorNotReady) return true;
            if (q == cudaSuccess) s.pending = false;   // complete: nothing to wait for any more
        }
    return false;
}
static int launch_blind_rotate(tfhe_b200_ctx* ctx, BrArgs& a, cudaStream_t st, bool timed) {
    a.bkdev = ctx->bkdev; a.mask = ctx->prm.decomp_mask; a.mu = ctx->prm.mu; a.ns = ctx->ns_int();
    if (a.split <= 0) a.split = a.B;
    const int slot = (int)(ctx->timed % tfhe_b200_ctx::RING);
    if (timed) CK(cudaEventRecord(ctx->ev[slot][0], st));
    // Launch shapes.  B <= #SMs: one gate per CTA, 254 registers (latency shape).  Larger batches: CTAs of up to G gates
    // (default G = 4: 24 warps = 6 per SM sub-partition, 80 registers, one CTA per SM -- measured 57.6 k gates/s against
    // 50.7 k for three 1-gate CTAs per SM whose 18 warps load the four sub-partitions 5/5/4/4); the batch is dealt out
    // evenly over rounds * #SMs CTAs so that a batch that is not a multiple of G * #SMs ends with 3-gate CTAs instead of
    // a half-empty last wave.
    const bool full = a.B > (long)ctx->sm_count;
    const int variant = (a.ns == 2 && ctx->variant != 9 && ctx->variant != 8) ? 7 : ctx->variant;   // the other measured alternatives exist for three slices only
    auto deal = [&](int G) {
        const long cap = (long)G * ctx->sm_count;
        const long rounds = (a.B + cap - 1) / cap;
        long nctas = rounds * ctx->sm_count;
        if (nctas > a.B) nctas = a.B;
        a.cta_base = (int)(a.B / nctas); a.cta_rem = (int)(a.B % nctas);
        ctx->gates_per_cta = G;
        return (unsigned)nctas;
    };
    auto fixed = [&](int G) {   // every CTA owns exactly G gates (the last one possibly fewer)
        const long nctas = (a.B + G - 1) / G;
        a.cta_base = (int)(a.B / nctas); a.cta_rem = (int)(a.B % nctas);
        ctx->gates_per_cta = G;
        return (unsigned)nctas;
    };
    // FFT64 mode: up to three gates per SM run as waves of the one-gate-per-SM latency kernel (2.4 ms a wave: 296 gates 4.8 ms against
    // 6.4 ms for K5T, 444 gates 7.2 ms against 7.7 ms for K5F with three gates per SM; four waves would take 9.5 ms against 7.9 ms)
    const bool f64_waves = ctx->key_slices == 1 && ctx->f64_latency && variant != 9 && variant != 8 && a.B <= 3L * ctx->sm_count;
    if (f64_waves) {
        // FFT64 latency shape, one gate per SM on twelve warps (two per transform): 2.37 ms per gate from 1 to #SMs gates, against
        // 2.66-2.73 ms for the NTT cluster kernel (two SMs per gate, at most #SMs/2 gates) and 3.57 ms for the one-CTA NTT kernel at
        // 148 gates.  TFHE_B200_F64_LATENCY=0: the NTT latency shapes.  (One warp per transform on six warps measured 3.2 ms.)
        a.cta_base = 1; a.cta_rem = 0;
        ctx->gates_per_cta = 1;
        if (ctx->f64_cluster && a.B <= (long)ctx->pair_max)   // one gate on a cluster of two SMs (K5FL3)
            blind_rotate_f64_latency3_kernel<<<(unsigned)(2 * a.B), F64L3_THREADS, F64L3_SMEM_BYTES, st>>>(a, ctx->bkdev_f64l);
        else
            blind_rotate_f64_latency2_kernel<<<(unsigned)a.B, F64L2_THREADS, F64L2_SMEM_BYTES, st>>>(a, ctx->bkdev_f64l);
    } else if (full && (variant == 3 || a.B <= 2L * ctx->sm_count) && a.ns == 3) {
        // 1-gate CTAs, up to three per SM (96 registers): the earlier default (TFHE_B200_BR_VARIANT=3 for A/B runs) and still the
        // best shape between one and two gates per SM (296 gates: 5.7 ms against 6.3 ms for 2-gate CTAs of the 80-register
        // kernel and 5.8 ms for 1-gate CTAs compiled for 168 registers)
        blind_rotate_kernel<1, false, 3><<<fixed(1), THREADS_PER_GATE, br_smem_bytes(1), st>>>(a);
    } else if (full && ctx->key_slices == 1 && variant != 8) {
        // FFT64 mode: one warp per gate, eight gates per CTA (blind_rotate_f64.cuh), above three gates per SM
        if (ctx->f64_tmem == 2) {
            const unsigned grid = batches_overlap(ctx, st) ? fixed(F64W2_GATES) : deal(F64W2_GATES);
            blind_rotate_f64w2_kernel<<<grid, F64W2_THREADS, f64w2_smem_bytes(), st>>>(a, ctx->bkdev_f64l);
        } else if (ctx->f64_tmem) {
            const unsigned grid = batches_overlap(ctx, st) ? fixed(F64T_GATES) : deal(F64T_GATES);
            blind_rotate_f64t_kernel<<<grid, F64T_GATES * 32, f64t_smem_bytes(), st>>>(a, ctx->bkdev_f64);
        } else {
            const unsigned grid = batches_overlap(ctx, st) ? fixed(F64_GATES) : deal(F64_GATES);
            blind_rotate_f64_kernel<<<grid, F64_GATES * 32, f64_smem_bytes(), st>>>(a, ctx->bkdev_f64, ctx->f64_stagger_ns);
        }
    } else if (full && a.ns == 2 && variant != 8) {   // default: the two-warps-per-gate throughput kernel (blind_rotate_t2.cuh)
        const int G = ctx->t2_gates;
        const unsigned grid = batches_overlap(ctx, st) ? fixed(G) : deal(G);
        a.bkdev = ctx->bkdev_t2;
        const size_t sm = t2_smem_bytes(G);
        const int tw = ctx->t2_twreg;
        if (G == 6) {
            if (tw == 0) blind_rotate_t2_kernel<6, 0><<<grid, 6 * T2_THREADS_PER_GATE, sm, st>>>(a);
            else if (tw == 3) blind_rotate_t2_kernel<6, 3><<<grid, 6 * T2_THREADS_PER_GATE, sm, st>>>(a);
            else blind_rotate_t2_kernel<6, 1><<<grid, 6 * T2_THREADS_PER_GATE, sm, st>>>(a);
        } else {
            if (tw == 0) blind_rotate_t2_kernel<4, 0><<<grid, 4 * T2_THREADS_PER_GATE, sm, st>>>(a);
            else if (tw == 3) blind_rotate_t2_kernel<4, 3><<<grid, 4 * T2_THREADS_PER_GATE, sm, st>>>(a);
            else blind_rotate_t2_kernel<4, 1><<<grid, 4 * T2_THREADS_PER_GATE, sm, st>>>(a);
        }
    } else if (full) {   // three slices (variant 7), or two slices on the six-warps-per-gate kernel (variant 8, for A/B runs)
        const unsigned grid = batches_overlap(ctx, st) ? fixed(4) : deal(4);
        if (a.ns == 2) blind_rotate_kernel<4, false, 1, 2><<<grid, 4 * THREADS_PER_GATE, br_smem_bytes(4), st>>>(a);
        else blind_rotate_kernel<4, false, 1, 3><<<grid, 4 * THREADS_PER_GATE, br_smem_bytes(4), st>>>(a);
    } else if (a.B <= (long)ctx->pair_max && variant != 9) {   // latency shape: one gate on a cluster of two SMs, as long as
                                                                // the clusters fit in one wave (74 gates: 3.66 ms against 3.92 ms
                                                                // with one CTA per gate; TFHE_B200_PAIR_MAX moves the limit)
        a.cta_base = 1; a.cta_rem = 0;
        ctx->gates_per_cta = 1;
        if (a.ns == 2) blind_rotate_pair_kernel<2><<<(unsigned)(2 * a.B), PAIR_LAUNCH_THREADS, (size_t)PAIR_SMEM_WORDS * 4, st>>>(a);
        else blind_rotate_pair_kernel<3><<<(unsigned)(2 * a.B), PAIR_LAUNCH_THREADS, (size_t)PAIR_SMEM_WORDS * 4, st>>>(a);
    } else if (ctx->slab_tma) {   // one gate per CTA, 254 registers, key slabs staged by bulk copies
        if (a.ns == 2) blind_rotate_kernel<1, false, 1, 2, true><<<fixed(1), THREADS_PER_GATE, br_smem_bytes_tma(), st>>>(a);
        else blind_rotate_kernel<1, false, 1, 3, true><<<fixed(1), THREADS_PER_GATE, br_smem_bytes_tma(), st>>>(a);
    } else {
        if (a.ns == 2) blind_rotate_kernel<1, false, 1, 2><<<fixed(1), THREADS_PER_GATE, br_smem_bytes(1), st>>>(a);
        else blind_rotate_kernel<1, false, 1><<<fixed(1), THREADS_PER_GATE, br_smem_bytes(1), st>>>(a);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    if (timed) CK(cudaEventRecord(ctx->ev[slot][1], st));
    return TFHE_B200_OK;
}
static int launch_keyswitch(tfhe_b200_ctx* ctx, const uint16_t* dig, uint32_t* out, long B, cudaStream_t st, bool timed,
                            const int32_t* idxo = nullptr) {
    const int slot = (int)(ctx->timed % tfhe_b200_ctx::RING);
    if (timed) CK(cudaEventRecord(ctx->ev[slot][2], st));
    if (ctx->ks_variant == 1) {   // register-tile kernel (TFHE_B200_KS_VARIANT=1)
        const long tile = (long)KS_GT * KS_GROUPS;
        const long tiles = (B + tile - 1) / tile;
        int isplit = KS_ISPLIT_MIN;   // small batches (latency path, narrow circuit levels): split the key indices further to fill the SMs
        while (isplit < 128 && tiles * isplit < 2L * ctx->sm_count) isplit *= 2;
        dim3 grid((unsigned)tiles, isplit);
        keyswitch_kernel<<<grid, dim3(KS_THREADS, KS_GROUPS), 0, st>>>(reinterpret_cast<const uint4*>(ctx->kskdev), dig, out, B, idxo);
    } else if (ctx->ks_variant == 3) {   // producer / consumer pipeline over the staged rows, one warp per gate (default)
        const long tiles = (B + KSP_GATES - 1) / KSP_GATES;
        // slices of the key indices: the count (8 .. 128) whose tiles x slices CTAs run in the fewest index-steps, two CTAs per SM:
        // whole waves matter -- 1024 gates = 64 tiles: 8 slices are 512 CTAs = 1.73 waves of 128 indices, 9 slices 1.95 waves of 114
        const long slots = 2L * ctx->sm_count;
        int isplit = KS_ISPLIT_MIN;
        long best = LONG_MAX;
        for (int y = KS_ISPLIT_MIN; y <= 128; y++) {
            const long waves = (tiles * y + slots - 1) / slots;
            const long cost = waves * ((1024 + y - 1) / y + 6);   // + 6: start-up of a CTA (digits, pipeline fill) in index-steps
            if (cost < best) { best = cost; isplit = y; }
        }
        keyswitch_p_kernel<KSP_GATES, KSP_RING><<<dim3((unsigned)tiles, isplit), (KSP_GATES + 1) * 32, KsP<KSP_GATES, KSP_RING>::SMEM_BYTES, st>>>(
            reinterpret_cast<const uint4*>(ctx->kskdev), dig, out, B, idxo);
    } else {                      // the same with one __syncthreads per stage instead of empty barriers (TFHE_B200_KS_VARIANT=2)
        const long tiles = (B + KS2_GATES - 1) / KS2_GATES;
        int isplit = KS_ISPLIT_MIN;
        while (isplit < 128 && tiles * isplit < 2L * ctx->sm_count) isplit *= 2;
        dim3 grid((unsigned)tiles, isplit);
        keyswitch2_kernel<<<grid, KS2_THREADS, KS2_SMEM_BYTES, st>>>(reinterpret_cast<const uint4*>(ctx->kskdev), dig, out, B, idxo);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    if (timed) { CK(cudaEventRecord(ctx->ev[slot][3], st)); ctx->timed++; }
    return TFHE_B200_OK;
}
static const size_t CT_WORDS = (size_t)(LWE_N + 1);
static const size_t CT_BYTES = CT_WORDS * 4;

// one bootstrapped gate batch on `st` with the workspaces of slot `s` (device pointers)
static int run_gates(tfhe_b200_ctx* ctx, Slot* s, BrArgs& a, uint32_t* out, cudaStream_t st) {
    RC(grow(ctx, (void**)&s->ksdig, &s->ksdig_cap, (size_t)a.B * 1024 * sizeof(uint16_t)));
    a.nsteps = LWE_N; a.out_init = out; a.ksdig = s->ksdig;
    RC(launch_blind_rotate(ctx, a, st, true));
    RC(launch_keyswitch(ctx, s->ksdig, out, a.B, st, true, a.idxo));
    ctx->last_batch = (uint64_t)a.B;
    return TFHE_B200_OK;
}
static int gate_args(tfhe_b200_ctx* ctx, const char* who, int op, const uint32_t* in0, const uint32_t* in1, size_t B, BrArgs* a) {
    if (op < 0 || op > TFHE_B200_ANDNY) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch: bad opcode");
    if (!ctx->have_bk || !ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "gate_batch: keys not loaded");
    bool two;
    op_coeffs(op, ctx->prm.mu, &a->c0, &a->c1, &a->cb, &two);
    if (two && !in1) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch: in1 required for this opcode");
    a->in0 = in0; a->in1 = two ? in1 : nullptr; a->B = (long)B;
    (void)who;
    return TFHE_B200_OK;
}
// hom_mux on device pointers: stage 1 = ONE launch of 2B gates (AND(control, in1) | AND(-control, in0)), stage 2 = OR
static int run_mux(tfhe_b200_ctx* ctx, Slot* s, const uint32_t* control, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B,
                   cudaStream_t st) {
    if (!ctx->have_bk || !ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "mux_batch: keys not loaded");
    RC(grow(ctx, (void**)&s->scratch, &s->scratch_cap, 2 * B * CT_BYTES));
    uint32_t* t = s->scratch;   // [0,B) = i_1 = hom_and(control, input_1); [B,2B) = i_0 = hom_and(-control, input_0)   (tfhe.rs:33-34)
    BrArgs a{};
    bool two;
    op_coeffs(TFHE_B200_AND, ctx->prm.mu, &a.c0, &a.c1, &a.cb, &two);
    op_coeffs(TFHE_B200_ANDNY, ctx->prm.mu, &a.c0b, &a.c1b, &a.cbb, &two);
    a.in0 = control; a.in1 = in1; a.in0b = control; a.in1b = in0; a.split = (long)B; a.B = (long)(2 * B);
    RC(run_gates(ctx, s, a, t, st));
    BrArgs b{};
    op_coeffs(TFHE_B200_OR, ctx->prm.mu, &b.c0, &b.c1, &b.cb, &two);   // bootstrap(i_1 + i_0 + 1/8)   (tfhe.rs:35-39)
    b.in0 = t; b.in1 = t + B * CT_WORDS; b.B = (long)B;
    return run_gates(ctx, s, b, out, st);
}

extern "C" {

int tfhe_b200_gate_batch_device(tfhe_b200_ctx* ctx, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B,
                                void* stream) {
    if (!ctx || !in0 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch: null argument");
    BrArgs a{};
    RC(gate_args(ctx, "gate_batch_device", op, in0, in1, B, &a));
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Slot* s;
    RC(slot_acquire(ctx, &st, false, &s));
    RC(run_gates(ctx, s, a, out, st));
    return slot_release(ctx, s, st);
}

// host pointers, asynchronous: H2D, gate batch, D2H are enqueued on one of the ctx's internal streams and the call
// returns; tfhe_b200_sync waits.  Host buffers should be pinned (pageable memory makes the copies synchronous).
int tfhe_b200_gate_batch_async(tfhe_b200_ctx* ctx, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B) {
    if (!ctx || !in0 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch: null argument");
    BrArgs a{};
    RC(gate_args(ctx, "gate_batch", op, in0, in1, B, &a));
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = nullptr;
    Slot* s;
    RC(slot_acquire(ctx, &st, true, &s));
    const size_t bytes = B * CT_BYTES;
    RC(grow(ctx, (void**)&s->tmp[0], &s->tmp_cap[0], bytes));
    RC(grow(ctx, (void**)&s->tmp[3], &s->tmp_cap[3], bytes));
    CK(cudaMemcpyAsync(s->tmp[0], in0, bytes, cudaMemcpyHostToDevice, st));
    a.in0 = s->tmp[0];
    if (a.in1) {
        RC(grow(ctx, (void**)&s->tmp[1], &s->tmp_cap[1], bytes));
        CK(cudaMemcpyAsync(s->tmp[1], in1, bytes, cudaMemcpyHostToDevice, st));
        a.in1 = s->tmp[1];
    }
    RC(run_gates(ctx, s, a, s->tmp[3], st));
    CK(cudaMemcpyAsync(out, s->tmp[3], bytes, cudaMemcpyDeviceToHost, st));
    return slot_release(ctx, s, st);
}
int tfhe_b200_gate_batch(tfhe_b200_ctx* ctx, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B) {
    RC(tfhe_b200_gate_batch_async(ctx, op, in0, in1, out, B));
    return tfhe_b200_sync(ctx);
}
// One launch for a batch whose gates have DIFFERENT opcodes (one level of a circuit): ops[g] in TFHE_B200_NAND..ANDNY;
// in1[g] is ignored for NOT / COPY gates (in1 may be NULL only if every gate is one of those).
int tfhe_b200_gate_batch_mixed_device(tfhe_b200_ctx* ctx, const uint8_t* ops_dev, const uint32_t* in0, const uint32_t* in1, uint32_t* out,
                                      size_t B, void* stream) {
    if (!ctx || !ops_dev || !in0 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch_mixed: null argument");
    if (!ctx->have_bk || !ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "gate_batch_mixed: keys not loaded");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Slot* s;
    RC(slot_acquire(ctx, &st, false, &s));
    BrArgs a{};
    a.ops = ops_dev; a.in0 = in0; a.in1 = in1; a.B = (long)B;
    RC(run_gates(ctx, s, a, out, st));
    return slot_release(ctx, s, st);
}
int tfhe_b200_gate_batch_mixed(tfhe_b200_ctx* ctx, const uint8_t* ops, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B) {
    if (!ctx || !ops || !in0 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch_mixed: null argument");
    if (!ctx->have_bk || !ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "gate_batch_mixed: keys not loaded");
    for (size_t g = 0; g < B; g++) {
        if (ops[g] > TFHE_B200_ANDNY) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch_mixed: bad opcode");
        if (!in1 && ops[g] != TFHE_B200_NOT && ops[g] != TFHE_B200_COPY) return fail(ctx, TFHE_B200_ERR_PARAM, "gate_batch_mixed: in1 required");
    }
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = nullptr;
    Slot* s;
    RC(slot_acquire(ctx, &st, true, &s));
    const size_t bytes = B * CT_BYTES;
    RC(grow(ctx, (void**)&s->tmp[0], &s->tmp_cap[0], bytes));
    RC(grow(ctx, (void**)&s->tmp[3], &s->tmp_cap[3], bytes));
    RC(grow(ctx, (void**)&s->opsbuf, &s->opsbuf_cap, B));
    CK(cudaMemcpyAsync(s->tmp[0], in0, bytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(s->opsbuf, ops, B, cudaMemcpyHostToDevice, st));
    BrArgs a{};
    a.ops = s->opsbuf; a.in0 = s->tmp[0]; a.B = (long)B;
    if (in1) {
        RC(grow(ctx, (void**)&s->tmp[1], &s->tmp_cap[1], bytes));
        CK(cudaMemcpyAsync(s->tmp[1], in1, bytes, cudaMemcpyHostToDevice, st));
        a.in1 = s->tmp[1];
    }
    RC(run_gates(ctx, s, a, s->tmp[3], st));
    CK(cudaMemcpyAsync(out, s->tmp[3], bytes, cudaMemcpyDeviceToHost, st));
    RC(slot_release(ctx, s, st));
    return tfhe_b200_sync(ctx);
}
int tfhe_b200_bootstrap_batch(tfhe_b200_ctx* ctx, const uint32_t* in, uint32_t* out, size_t B) {
    return tfhe_b200_gate_batch(ctx, TFHE_B200_COPY, in, nullptr, out, B);
}

int tfhe_b200_mux_batch_device(tfhe_b200_ctx* ctx, const uint32_t* control, const uint32_t* in0, const uint32_t* in1, uint32_t* out,
                               size_t B, void* stream) {
    if (!ctx || !control || !in0 || !in1 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "mux_batch: null argument");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Slot* s;
    RC(slot_acquire(ctx, &st, false, &s));
    RC(run_mux(ctx, s, control, in0, in1, out, B, st));
    return slot_release(ctx, s, st);
}
int tfhe_b200_mux_batch(tfhe_b200_ctx* ctx, const uint32_t* control, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B) {
    if (!ctx || !control || !in0 || !in1 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "mux_batch: null argument");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = nullptr;
    Slot* s;
    RC(slot_acquire(ctx, &st, true, &s));
    const size_t bytes = B * CT_BYTES;
    const uint32_t* src[3] = {control, in0, in1};
    for (int k = 0; k < 3; k++) {
        RC(grow(ctx, (void**)&s->tmp[k], &s->tmp_cap[k], bytes));
        CK(cudaMemcpyAsync(s->tmp[k], src[k], bytes, cudaMemcpyHostToDevice, st));
    }
    RC(grow(ctx, (void**)&s->tmp[3], &s->tmp_cap[3], bytes));
    RC(run_mux(ctx, s, s->tmp[0], s->tmp[1], s->tmp[2], s->tmp[3], B, st));
    CK(cudaMemcpyAsync(out, s->tmp[3], bytes, cudaMemcpyDeviceToHost, st));
    RC(slot_release(ctx, s, st));
    return tfhe_b200_sync(ctx);
}

}  // extern "C"

// ---- device-resident circuits: a levelised netlist uploaded once, evaluated with one launch pair per level and no host
// round trip (the reference walks the expression tree one gate at a time, nander/src/lib.rs:72-89) ----
struct tfhe_b200_circuit {
    std::vector<size_t> level_first, level_gates;
    uint8_t* ops = nullptr;       // device, all levels concatenated
    int32_t* idx = nullptr;       // device: [3][total] = in0 | in1 | out wire indices
    size_t total = 0, max_level = 0, n_wires = 0;
};
extern "C" {
int tfhe_b200_circuit_create(tfhe_b200_ctx* ctx, size_t n_levels, const size_t* level_gates, const uint8_t* ops, const int32_t* in0,
                             const int32_t* in1, const int32_t* out, size_t n_wires, tfhe_b200_circuit** circuit) {
    if (!ctx || !circuit || (n_levels && (!level_gates || !ops || !in0 || !in1 || !out)))
        return fail(ctx, TFHE_B200_ERR_PARAM, "circuit_create: null argument");
    *circuit = nullptr;
    tfhe_b200_circuit* c = new tfhe_b200_circuit();
    c->n_wires = n_wires;
    for (size_t l = 0; l < n_levels; l++) {
        c->level_first.push_back(c->total);
        c->level_gates.push_back(level_gates[l]);
        c->total += level_gates[l];
        if (level_gates[l] > c->max_level) c->max_level = level_gates[l];
    }
    for (size_t g = 0; g < c->total; g++) {
        const bool one = ops[g] == TFHE_B200_NOT || ops[g] == TFHE_B200_COPY;
        if (ops[g] > TFHE_B200_ANDNY || in0[g] < 0 || (size_t)in0[g] >= n_wires || out[g] < 0 || (size_t)out[g] >= n_wires ||
            (!one && (in1[g] < 0 || (size_t)in1[g] >= n_wires))) {
            delete c;
            return fail(ctx, TFHE_B200_ERR_PARAM, "circuit_create: opcode or wire index out of range");
        }
    }
    // A level is evaluated in place on one wire table by gates that run concurrently: within a level no wire may be written
    // twice, and no wire may be both read and written (the reader could see either value).  Sizes must fit the 32-bit grid math.
    if (n_wires > (size_t)INT32_MAX || c->max_level > ((size_t)1 << 30)) {
        delete c;
        return fail(ctx, TFHE_B200_ERR_PARAM, "circuit_create: more than 2^31 - 1 wires or 2^30 gates in a level");
    }
    {
        std::vector<uint32_t> written(n_wires, 0u), read(n_wires, 0u);   // level number + 1 of the last write / read
        size_t g = 0;
        for (size_t l = 0; l < n_levels; l++) {
            const uint32_t tag = (uint32_t)l + 1u;
            for (size_t k = 0; k < level_gates[l]; k++, g++) {
                const bool one = ops[g] == TFHE_B200_NOT || ops[g] == TFHE_B200_COPY;
                read[in0[g]] = tag;
                if (!one) read[in1[g]] = tag;
            }
            g -= level_gates[l];
            for (size_t k = 0; k < level_gates[l]; k++, g++) {
                if (written[out[g]] == tag || read[out[g]] == tag) {
                    delete c;
                    return fail(ctx, TFHE_B200_ERR_PARAM, "circuit_create: a wire is written twice, or read and written, within one level");
                }
                written[out[g]] = tag;
            }
        }
    }
    if (c->total) {
        cudaError_t e = cudaSetDevice(ctx->device);
        if (e == cudaSuccess) e = cudaMalloc(&c->ops, c->total);
        if (e == cudaSuccess) e = cudaMalloc(&c->idx, 3 * c->total * sizeof(int32_t));
        if (e == cudaSuccess) e = cudaMemcpy(c->ops, ops, c->total, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(c->idx, in0, c->total * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(c->idx + c->total, in1, c->total * 4, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(c->idx + 2 * c->total, out, c->total * 4, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            ctx->err = std::string("circuit_create: ") + cudaGetErrorString(e);
            cudaFree(c->ops); cudaFree(c->idx);
            delete c;
            return TFHE_B200_ERR_CUDA;
        }
    }
    *circuit = c;
    return TFHE_B200_OK;
}
int tfhe_b200_circuit_destroy(tfhe_b200_ctx* ctx, tfhe_b200_circuit* c) {
    if (!ctx || !c) return TFHE_B200_ERR_PARAM;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(c->ops); cudaFree(c->idx);
    delete c;
    return TFHE_B200_OK;
}
// wires_dev: [n_wires][n+1] on the device; inputs and constants filled in by the caller, gate outputs written in place.
// Everything is enqueued on `stream`; levels are ordered by the stream, the gates of a level are one blind-rotate + one
// key-switch launch.
int tfhe_b200_circuit_run_device(tfhe_b200_ctx* ctx, const tfhe_b200_circuit* c, uint32_t* wires_dev, void* stream) {
    if (!ctx || !c || !wires_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "circuit_run: null argument");
    if (!ctx->have_bk || !ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "circuit_run: keys not loaded");
    if (c->total == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Slot* s;
    RC(slot_acquire(ctx, &st, false, &s));
    RC(grow(ctx, (void**)&s->ksdig, &s->ksdig_cap, c->max_level * 1024 * sizeof(uint16_t)));
    for (size_t l = 0; l < c->level_gates.size(); l++) {
        const size_t f = c->level_first[l], n = c->level_gates[l];
        if (n == 0) continue;
        BrArgs a{};
        a.ops = c->ops + f;
        a.idx0 = c->idx + f; a.idx1 = c->idx + c->total + f; a.idxo = c->idx + 2 * c->total + f;
        a.in0 = wires_dev; a.in1 = wires_dev; a.B = (long)n;
        RC(run_gates(ctx, s, a, wires_dev, st));
    }
    return slot_release(ctx, s, st);
}
// ---- pieces of a run for evaluators that spread a level over several devices (group.cu, SURVEY 8e: "per level, all-gather of
// that level's outputs so every GPU holds all wires") ----
int tfhe_b200_circuit_shape(const tfhe_b200_circuit* c, size_t* n_levels, size_t* n_wires, size_t* max_level_gates) {
    if (!c) return TFHE_B200_ERR_PARAM;
    if (n_levels) *n_levels = c->level_gates.size();
    if (n_wires) *n_wires = c->n_wires;
    if (max_level_gates) *max_level_gates = c->max_level;
    return TFHE_B200_OK;
}
int tfhe_b200_circuit_level_gates(const tfhe_b200_circuit* c, size_t level, size_t* gates) {
    if (!c || !gates || level >= c->level_gates.size()) return TFHE_B200_ERR_PARAM;
    *gates = c->level_gates[level];
    return TFHE_B200_OK;
}
// gates [first, first + count) of one level.  rows_out == NULL: results go to their wires (as circuit_run_device does);
// otherwise row k of rows_out receives the output of gate first + k and the wire table is only read.
int tfhe_b200_circuit_run_level_device(tfhe_b200_ctx* ctx, const tfhe_b200_circuit* c, size_t level, size_t first, size_t count,
                                       uint32_t* wires_dev, uint32_t* rows_out, void* stream) {
    if (!ctx || !c || !wires_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "circuit_run_level: null argument");
    if (!ctx->have_bk || !ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "circuit_run_level: keys not loaded");
    if (level >= c->level_gates.size() || first > c->level_gates[level] || count > c->level_gates[level] - first)
        return fail(ctx, TFHE_B200_ERR_PARAM, "circuit_run_level: level or gate range out of bounds");
    if (count == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Slot* s;
    RC(slot_acquire(ctx, &st, false, &s));
    // consecutive calls rotate over the work slots: size each slot's digit buffer for the widest level once, not level by level
    RC(grow(ctx, (void**)&s->ksdig, &s->ksdig_cap, c->max_level * 1024 * sizeof(uint16_t)));
    const size_t f = c->level_first[level] + first;
    BrArgs a{};
    a.ops = c->ops + f;
    a.idx0 = c->idx + f; a.idx1 = c->idx + c->total + f; a.idxo = rows_out ? nullptr : c->idx + 2 * c->total + f;
    a.in0 = wires_dev; a.in1 = wires_dev; a.B = (long)count;
    RC(run_gates(ctx, s, a, rows_out ? rows_out : wires_dev, st));
    return slot_release(ctx, s, st);
}
// wires[out wire of gate k of the level] = rows[k] for every gate of the level
int tfhe_b200_circuit_scatter_level_device(tfhe_b200_ctx* ctx, const tfhe_b200_circuit* c, size_t level, const uint32_t* rows_dev,
                                           uint32_t* wires_dev, void* stream) {
    if (!ctx || !c || !rows_dev || !wires_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "circuit_scatter_level: null argument");
    if (level >= c->level_gates.size()) return fail(ctx, TFHE_B200_ERR_PARAM, "circuit_scatter_level: level out of bounds");
    const size_t n = c->level_gates[level];
    if (n == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    wire_scatter_kernel<<<(unsigned)((n + 7) / 8), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(rows_dev),
                                                                                 c->idx + 2 * c->total + c->level_first[level], wires_dev, (long)n);
    ctx->launches++;
    CK(cudaGetLastError());
    return TFHE_B200_OK;
}
}  // extern "C"

// ---- device-side key generation, encryption, decryption ----
extern "C" {

static int keygen_device_impl(tfhe_b200_ctx* ctx, const tfhe_rng::RngKey& seed, const uint8_t* s0, const uint8_t* s1) {
    if (!ctx || !s0 || !s1) return fail(ctx, TFHE_B200_ERR_PARAM, "keygen_device: null argument");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->slots[0].stream;
    uint8_t hb[2048];
    int32_t hp[1024];
    memset(hb, 0, sizeof hb);
    memcpy(hb, s0, LWE_N);
    memcpy(hb + 1024, s1, 1024);
    for (int k = 0; k < 1024; k++) hp[k] = s1[k] ? 1 : 0;
    CK(cudaMemcpyAsync(ctx->keybits, hb, sizeof hb, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ctx->s1poly, hp, sizeof hp, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));   // hb / hp live on this stack frame
    const int rows = LWE_N * 6;
    // BK rows: A uniform, B = noise; B += A * s1 (exact negacyclic product); gadget term; NTT-domain transform
    bk_fill_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->bk_torus, seed, (long)rows * 1024);
    polymul_kernel<<<(rows + PM_WARPS - 1) / PM_WARPS, PM_WARPS * 32, 0, st>>>(ctx->bk_torus + 1024, ctx->s1poly, ctx->bk_torus, rows, 2048, 0,
                                                                                 2048, 1);
    bk_gadget_kernel<<<(rows + 255) / 256, 256, 0, st>>>(ctx->bk_torus, ctx->keybits, rows);
    ctx->launches += 3;
    CK(cudaGetLastError());
    RC(transform_keys(ctx, ctx->bk_torus, ctx->bkdev, LWE_N, st));
    // KSK rows straight into the device key
    const long krows = 1024L * 8 * 3;
    lwe_rows_kernel<<<(unsigned)((krows + 7) / 8), 256, 0, st>>>(ctx->kskdev, krows, seed, 0, ctx->keybits, ctx->keybits + 1024, nullptr, 0);
    ctx->launches++;
    CK(cudaGetLastError());
    // the secret keys do not outlive the key generation on the device
    CK(cudaMemsetAsync(ctx->keybits, 0, 2048, st));
    CK(cudaMemsetAsync(ctx->s1poly, 0, 1024 * 4, st));
    CK(cudaStreamSynchronize(st));
    ctx->have_bk = ctx->have_ksk = true;
    return TFHE_B200_OK;
}
// generator key of a *_csprng call: the caller's 32 bytes, or fresh OS entropy (getrandom) when key == NULL
static int csprng_key(tfhe_b200_ctx* ctx, const uint8_t* key, tfhe_rng::RngKey* out) {
    uint8_t fresh[32];
    if (!key) {
        if (tfhe_b200_random_bytes(fresh, sizeof fresh) != TFHE_B200_OK) return fail(ctx, TFHE_B200_ERR_IO, "getrandom failed");
        key = fresh;
    }
    *out = tfhe_rng::key_from_bytes(key);
    volatile uint8_t* w = fresh;
    for (size_t i = 0; i < sizeof fresh; i++) w[i] = 0;
    return TFHE_B200_OK;
}
int tfhe_b200_keygen_device(tfhe_b200_ctx* ctx, uint64_t seed, const uint8_t* s0, const uint8_t* s1) {
    return keygen_device_impl(ctx, tfhe_rng::key_from_seed(seed), s0, s1);
}
int tfhe_b200_keygen_device_csprng(tfhe_b200_ctx* ctx, const uint8_t* key, const uint8_t* s0, const uint8_t* s1) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    tfhe_rng::RngKey k;
    RC(csprng_key(ctx, key, &k));
    return keygen_device_impl(ctx, k, s0, s1);
}
int tfhe_b200_export_bk(tfhe_b200_ctx* ctx, uint32_t* bk_host) {
    if (!ctx || !bk_host) return fail(ctx, TFHE_B200_ERR_PARAM, "export_bk: null argument");
    if (!ctx->have_bk) return fail(ctx, TFHE_B200_ERR_STATE, "export_bk: bootstrapping key not loaded");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(bk_host, ctx->bk_torus, BK_TORUS_BYTES, cudaMemcpyDeviceToHost));
    return TFHE_B200_OK;
}
// device-to-device forms: the source of the one-off NCCL broadcast that replicates the keys to the other GPUs
int tfhe_b200_export_bk_device(tfhe_b200_ctx* ctx, uint32_t* bk_dev, void* stream) {
    if (!ctx || !bk_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "export_bk_device: null argument");
    if (!ctx->have_bk) return fail(ctx, TFHE_B200_ERR_STATE, "export_bk_device: bootstrapping key not loaded");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(bk_dev, ctx->bk_torus, BK_TORUS_BYTES, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return TFHE_B200_OK;
}
int tfhe_b200_export_ksk_device(tfhe_b200_ctx* ctx, uint32_t* ksk_dev, void* stream) {
    if (!ctx || !ksk_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "export_ksk_device: null argument");
    if (!ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "export_ksk_device: key-switching key not loaded");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ksk_dev, ctx->kskdev, KSK_BYTES, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return TFHE_B200_OK;
}
int tfhe_b200_export_ksk(tfhe_b200_ctx* ctx, uint32_t* ksk_host) {
    if (!ctx || !ksk_host) return fail(ctx, TFHE_B200_ERR_PARAM, "export_ksk: null argument");
    if (!ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "export_ksk: key-switching key not loaded");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(ksk_host, ctx->kskdev, KSK_BYTES, cudaMemcpyDeviceToHost));
    return TFHE_B200_OK;
}
// bits_dev: [B] bytes on the device; out_dev: [B][n+1] on the device; s0: host
static int encrypt_bits_device_impl(tfhe_b200_ctx* ctx, const tfhe_rng::RngKey& seed, uint64_t ct_index0, const uint8_t* s0, const uint8_t* bits_dev,
                                    size_t B, uint32_t* out_dev, void* stream) {
    if (!ctx || !s0 || (B && (!bits_dev || !out_dev))) return fail(ctx, TFHE_B200_ERR_PARAM, "encrypt_bits_device: null argument");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Slot* s;
    RC(slot_acquire(ctx, &st, false, &s));
    CK(cudaMemcpyAsync(s->s0buf, s0, LWE_N, cudaMemcpyHostToDevice, st));
    lwe_rows_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(out_dev, (long)B, seed, ct_index0, s->s0buf, nullptr, bits_dev, 1);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(s->s0buf, 0, 1024, st));   // the secret key does not stay on the device
    return slot_release(ctx, s, st);
}
int tfhe_b200_encrypt_bits_device(tfhe_b200_ctx* ctx, uint64_t seed, uint64_t ct_index0, const uint8_t* s0, const uint8_t* bits_dev,
                                  size_t B, uint32_t* out_dev, void* stream) {
    return encrypt_bits_device_impl(ctx, tfhe_rng::key_from_seed(seed), ct_index0, s0, bits_dev, B, out_dev, stream);
}
int tfhe_b200_encrypt_bits_device_csprng(tfhe_b200_ctx* ctx, const uint8_t* key, const uint8_t* s0, const uint8_t* bits_dev, size_t B,
                                         uint32_t* out_dev, void* stream) {
    if (!ctx) return TFHE_B200_ERR_PARAM;
    tfhe_rng::RngKey k;
    RC(csprng_key(ctx, key, &k));
    return encrypt_bits_device_impl(ctx, k, 0, s0, bits_dev, B, out_dev, stream);
}
// phase_dev / bits_dev: either may be NULL
int tfhe_b200_decrypt_bits_device(tfhe_b200_ctx* ctx, const uint8_t* s0, const uint32_t* ct_dev, size_t B, uint8_t* bits_dev,
                                  uint32_t* phase_dev, void* stream) {
    if (!ctx || !s0 || (B && !ct_dev)) return fail(ctx, TFHE_B200_ERR_PARAM, "decrypt_bits_device: null argument");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Slot* s;
    RC(slot_acquire(ctx, &st, false, &s));
    CK(cudaMemcpyAsync(s->s0buf, s0, LWE_N, cudaMemcpyHostToDevice, st));
    lwe_phase_kernel<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(ct_dev, (long)B, s->s0buf, phase_dev, bits_dev);
    ctx->launches++;
    CK(cudaGetLastError());
    CK(cudaMemsetAsync(s->s0buf, 0, 1024, st));   // the secret key does not stay on the device
    return slot_release(ctx, s, st);
}

}  // extern "C"

// external product / cmux on device pointers: transform the TRGSW samples into slot scratch, then one CMUX step per product
static int run_extprod(tfhe_b200_ctx* ctx, Slot* s, const uint32_t* trgsw_dev, size_t ntrgsw, const uint32_t* rep1, const uint32_t* rep0,
                       uint32_t* out, size_t B, cudaStream_t st) {
    BrArgs a{};
    a.bkdev = s->scratch; a.mask = ctx->prm.decomp_mask; a.mu = ctx->prm.mu; a.B = (long)B; a.split = (long)B; a.nsteps = 1;
    a.trlwe_in = rep1; a.trlwe_in0 = rep0; a.trlwe_out = out; a.ntrgsw = (long)ntrgsw; a.ns = ctx->ns_int();
    const bool rows16 = (((uintptr_t)rep1 | (uintptr_t)rep0 | (uintptr_t)out) & 15) == 0;   // the FFT64 kernel moves its rows as 16-byte words
    if (ctx->key_slices == 1 && ntrgsw == 1 && B > (size_t)ctx->sm_count && rows16) {   // FFT64, one shared TRGSW: persistent CTAs, one product per warp
        RC(grow(ctx, (void**)&s->scratch, &s->scratch_cap, (size_t)BK_STEP_WORDS * 4));
        cd16* kx = reinterpret_cast<cd16*>(s->scratch);   // 96 KB of the slot scratch
        bk_transform_f64_kernel<<<(12 + KTF_WARPS - 1) / KTF_WARPS, KTF_WARPS * 32, 0, st>>>(trgsw_dev, kx, 12);
        const long ngroups = ((long)B + F64_GATES - 1) / F64_GATES;
        const unsigned grid = (unsigned)std::min<long>(ngroups, (long)ctx->sm_count);
        external_product_f64_kernel<<<grid, F64_GATES * 32, f64_smem_bytes(), st>>>(a, kx);
        ctx->launches += 2;
        CK(cudaGetLastError());
        return TFHE_B200_OK;
    }
    if (ctx->key_slices == 1 && ntrgsw == B && B > (size_t)ctx->sm_count && rows16) {   // FFT64, one TRGSW per item: the warp transforms it itself
        const long nct = ((long)B + XPI_WARPS - 1) / XPI_WARPS;
        const unsigned grid = (unsigned)std::min<long>(nct, (long)ctx->sm_count);
        external_product_item_f64_kernel<<<grid, XPI_WARPS * 32, XPI_SMEM_BYTES, st>>>(a, trgsw_dev);
        ctx->launches++;
        CK(cudaGetLastError());
        return TFHE_B200_OK;
    }
    RC(grow(ctx, (void**)&s->scratch, &s->scratch_cap, ntrgsw * BK_STEP_WORDS * 4));   // the NTT kernels read transformed TRGSWs from the slot scratch
    a.bkdev = s->scratch;
    if (a.ns == 2 && B > (size_t)ctx->sm_count) {   // two slices, throughput shape: the two-warps-per-product kernel, six products per CTA
        const int npolys = (int)ntrgsw * 12;
        bk_transform_t2_kernel<<<(npolys + KT_WARPS - 1) / KT_WARPS, KT_WARPS * 32, 0, st>>>(trgsw_dev, s->scratch, npolys);
        const unsigned grid = (unsigned)((B + 5) / 6);
        a.cta_base = (int)(B / grid); a.cta_rem = (int)(B % grid);
        blind_rotate_t2_kernel<6, 1, true><<<grid, 6 * T2_THREADS_PER_GATE, t2_smem_bytes(6), st>>>(a);
        ctx->launches += 2;
        CK(cudaGetLastError());
        return TFHE_B200_OK;
    }
    RC(transform_keys(ctx, trgsw_dev, s->scratch, (int)ntrgsw, st));
    if (B > (size_t)ctx->sm_count) {   // throughput shape: 4 products per CTA
        const unsigned grid = (unsigned)((B + 3) / 4);
        a.cta_base = (int)(B / grid); a.cta_rem = (int)(B % grid);
        if (a.ns == 2) blind_rotate_kernel<4, true, 1, 2><<<grid, 4 * THREADS_PER_GATE, br_smem_bytes(4), st>>>(a);
        else blind_rotate_kernel<4, true, 1><<<grid, 4 * THREADS_PER_GATE, br_smem_bytes(4), st>>>(a);
    } else {
        const unsigned grid = (unsigned)((B + 1) / 2);
        a.cta_base = (int)(B / grid); a.cta_rem = (int)(B % grid);
        if (a.ns == 2) blind_rotate_kernel<2, true, 1, 2><<<grid, 2 * THREADS_PER_GATE, br_smem_bytes(2), st>>>(a);
        else blind_rotate_kernel<2, true, 1><<<grid, 2 * THREADS_PER_GATE, br_smem_bytes(2), st>>>(a);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return TFHE_B200_OK;
}

// ---- step-level entries: host pointers, synchronous.  `fn` enqueues the device work on (slot, stream). ----
struct HostIo {
    const void* in[3] = {nullptr, nullptr, nullptr};
    size_t in_bytes[3] = {0, 0, 0};
    void* out = nullptr;
    size_t out_bytes = 0;
};
template <class F>
static int with_host_io(tfhe_b200_ctx* ctx, const HostIo& io, F fn) {
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = nullptr;
    Slot* s;
    RC(slot_acquire(ctx, &st, true, &s));
    uint32_t* dev_in[3] = {nullptr, nullptr, nullptr};
    for (int k = 0; k < 3; k++) {
        if (!io.in[k]) continue;
        RC(grow(ctx, (void**)&s->tmp[k], &s->tmp_cap[k], io.in_bytes[k]));
        CK(cudaMemcpyAsync(s->tmp[k], io.in[k], io.in_bytes[k], cudaMemcpyHostToDevice, st));
        dev_in[k] = s->tmp[k];
    }
    RC(grow(ctx, (void**)&s->tmp[3], &s->tmp_cap[3], io.out_bytes));
    RC(fn(s, st, dev_in, s->tmp[3]));
    CK(cudaMemcpyAsync(io.out, s->tmp[3], io.out_bytes, cudaMemcpyDeviceToHost, st));
    RC(slot_release(ctx, s, st));
    return tfhe_b200_sync(ctx);
}

extern "C" {

int tfhe_b200_blind_rotate_batch(tfhe_b200_ctx* ctx, const uint32_t* in, int nsteps, uint32_t* out_trlwe, size_t B) {
    if (!ctx || !in || !out_trlwe) return fail(ctx, TFHE_B200_ERR_PARAM, "blind_rotate_batch: null argument");
    if (nsteps < 0 || nsteps > LWE_N) return fail(ctx, TFHE_B200_ERR_PARAM, "blind_rotate_batch: nsteps out of range");
    if (!ctx->have_bk) return fail(ctx, TFHE_B200_ERR_STATE, "blind_rotate_batch: bootstrapping key not loaded");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = in; io.in_bytes[0] = B * CT_BYTES; io.out = out_trlwe; io.out_bytes = B * 2048 * 4;
    return with_host_io(ctx, io, [&](Slot*, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        BrArgs a{};
        a.c0 = 1; a.in0 = di[0]; a.nsteps = nsteps; a.B = (long)B; a.trlwe_out = dout;
        return launch_blind_rotate(ctx, a, st, false);
    });
}
int tfhe_b200_bootstrap_lv1_batch(tfhe_b200_ctx* ctx, const uint32_t* in, uint32_t* out_lwe1, size_t B) {
    if (!ctx || !in || !out_lwe1) return fail(ctx, TFHE_B200_ERR_PARAM, "bootstrap_lv1_batch: null argument");
    if (!ctx->have_bk) return fail(ctx, TFHE_B200_ERR_STATE, "bootstrap_lv1_batch: bootstrapping key not loaded");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = in; io.in_bytes[0] = B * CT_BYTES; io.out = out_lwe1; io.out_bytes = B * 1025 * 4;
    return with_host_io(ctx, io, [&](Slot*, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        BrArgs a{};
        a.c0 = 1; a.in0 = di[0]; a.nsteps = LWE_N; a.B = (long)B; a.lwe1_out = dout;
        return launch_blind_rotate(ctx, a, st, false);
    });
}
int tfhe_b200_keyswitch_batch(tfhe_b200_ctx* ctx, const uint32_t* lwe1, uint32_t* out, size_t B) {
    if (!ctx || !lwe1 || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "keyswitch_batch: null argument");
    if (!ctx->have_ksk) return fail(ctx, TFHE_B200_ERR_STATE, "keyswitch_batch: key-switching key not loaded");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = lwe1; io.in_bytes[0] = B * 1025 * 4; io.out = out; io.out_bytes = B * CT_BYTES;
    return with_host_io(ctx, io, [&](Slot* s, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        RC(grow(ctx, (void**)&s->ksdig, &s->ksdig_cap, B * 1024 * sizeof(uint16_t)));
        lwe1_prepare_kernel<<<(unsigned)B, 256, 0, st>>>(di[0], s->ksdig, dout, (long)B);
        ctx->launches++;
        return launch_keyswitch(ctx, s->ksdig, dout, (long)B, st, false);
    });
}
int tfhe_b200_external_product_batch(tfhe_b200_ctx* ctx, const uint32_t* trgsw, size_t ntrgsw, const uint32_t* trlwe, uint32_t* out,
                                     size_t B) {
    if (!ctx || !trgsw || !trlwe || !out || ntrgsw == 0) return fail(ctx, TFHE_B200_ERR_PARAM, "external_product_batch: bad argument");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = trlwe; io.in_bytes[0] = B * 2048 * 4; io.in[1] = trgsw; io.in_bytes[1] = ntrgsw * 12 * 1024 * 4;
    io.out = out; io.out_bytes = B * 2048 * 4;
    return with_host_io(ctx, io, [&](Slot* s, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        return run_extprod(ctx, s, di[1], ntrgsw, di[0], nullptr, dout, B, st);
    });
}
// device-pointer forms of the two micro-benchmark entries (BASELINE config 3 measured device resident)
int tfhe_b200_external_product_batch_device(tfhe_b200_ctx* ctx, const uint32_t* trgsw_dev, size_t ntrgsw, const uint32_t* trlwe_dev,
                                            uint32_t* out_dev, size_t B, void* stream) {
    if (!ctx || !trgsw_dev || !trlwe_dev || !out_dev || ntrgsw == 0) return fail(ctx, TFHE_B200_ERR_PARAM, "external_product_batch_device: bad argument");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    Slot* s;
    RC(slot_acquire(ctx, &st, false, &s));
    RC(run_extprod(ctx, s, trgsw_dev, ntrgsw, trlwe_dev, nullptr, out_dev, B, st));
    return slot_release(ctx, s, st);
}
// exact negacyclic products: FFT64 mode and more than a wave of products -> polymul_f64_kernel, otherwise the NTT kernel
static int launch_polymul(tfhe_b200_ctx* ctx, const uint32_t* a_dev, const int32_t* d_dev, uint32_t* out_dev, size_t B, cudaStream_t st) {
    if (ctx->key_slices == 1 && B > (size_t)ctx->sm_count) {
        const long nctas = std::min<long>(((long)B + PMF_WARPS - 1) / PMF_WARPS, (long)ctx->sm_count);
        polymul_f64_kernel<<<(unsigned)nctas, PMF_WARPS * 32, PMF_SMEM_BYTES, st>>>(a_dev, d_dev, out_dev, (long)B);
    } else {
        polymul_kernel<<<(unsigned)((B + PM_WARPS - 1) / PM_WARPS), PM_WARPS * 32, 0, st>>>(a_dev, d_dev, out_dev, (long)B, 1024, 1024, 1024, 0);
    }
    ctx->launches++;
    CK(cudaGetLastError());
    return TFHE_B200_OK;
}
int tfhe_b200_negacyclic_mul_batch_device(tfhe_b200_ctx* ctx, const uint32_t* a_dev, const int32_t* d_dev, uint32_t* out_dev, size_t B,
                                          void* stream) {
    if (!ctx || !a_dev || !d_dev || !out_dev) return fail(ctx, TFHE_B200_ERR_PARAM, "negacyclic_mul_batch_device: null argument");
    if (B == 0) return TFHE_B200_OK;
    CK(cudaSetDevice(ctx->device));
    return launch_polymul(ctx, a_dev, d_dev, out_dev, B, (cudaStream_t)stream);
}
int tfhe_b200_negacyclic_mul_batch(tfhe_b200_ctx* ctx, const uint32_t* a, const int32_t* d, uint32_t* out, size_t B) {
    if (!ctx || !a || !d || !out) return fail(ctx, TFHE_B200_ERR_PARAM, "negacyclic_mul_batch: null argument");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = a; io.in_bytes[0] = B * 4096; io.in[1] = d; io.in_bytes[1] = B * 4096; io.out = out; io.out_bytes = B * 4096;
    return with_host_io(ctx, io, [&](Slot*, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        return launch_polymul(ctx, di[0], (const int32_t*)di[1], dout, B, st);
    });
}

// TRGSWRep::cmux(rep_1, rep_0) = cross(rep_1 - rep_0) + rep_0  (trgsw.rs:315-322, 323-330); trgsw in the torus domain
int tfhe_b200_cmux_batch(tfhe_b200_ctx* ctx, const uint32_t* trgsw, size_t ntrgsw, const uint32_t* rep1, const uint32_t* rep0,
                         uint32_t* out, size_t B) {
    if (!ctx || !trgsw || !rep1 || !rep0 || !out || ntrgsw == 0) return fail(ctx, TFHE_B200_ERR_PARAM, "cmux_batch: bad argument");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = rep1; io.in_bytes[0] = B * 2048 * 4; io.in[1] = trgsw; io.in_bytes[1] = ntrgsw * 12 * 1024 * 4;
    io.in[2] = rep0; io.in_bytes[2] = B * 2048 * 4; io.out = out; io.out_bytes = B * 2048 * 4;
    return with_host_io(ctx, io, [&](Slot* s, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        return run_extprod(ctx, s, di[1], ntrgsw, di[0], di[2], dout, B, st);
    });
}
// TRLWERep::sample_extract_index(index) (trlwe.rs:110-121): [B][2][N] -> [B][N+1]
int tfhe_b200_sample_extract_batch(tfhe_b200_ctx* ctx, const uint32_t* trlwe, int index, uint32_t* out_lwe1, size_t B) {
    if (!ctx || !trlwe || !out_lwe1) return fail(ctx, TFHE_B200_ERR_PARAM, "sample_extract_batch: null argument");
    if (index < 0 || index >= 1024) return fail(ctx, TFHE_B200_ERR_PARAM, "sample_extract_batch: index out of range");
    if (B == 0) return TFHE_B200_OK;
    HostIo io; io.in[0] = trlwe; io.in_bytes[0] = B * 2048 * 4; io.out = out_lwe1; io.out_bytes = B * 1025 * 4;
    return with_host_io(ctx, io, [&](Slot*, cudaStream_t st, uint32_t* const* di, uint32_t* dout) {
        sample_extract_kernel<<<(unsigned)B, 256, 0, st>>>(di[0], dout, (long)B, index);
        ctx->launches++;
        CK(cudaGetLastError());
        return TFHE_B200_OK;
    });
}

}  // extern "C"

