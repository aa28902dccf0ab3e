// group.cu -- multi-GPU behind the C ABI: one PROCESS drives several B200s of a box (include/tfhe_b200.h, "device groups").
//
// What the reference offers here: nothing -- `TFHE::new` (hom_nand/src/tfhe.rs:21-25) builds one key pair in host memory and
// every gate runs on the calling thread.  What north_star asks for: independent gates of a batch sharded over the GPUs of a
// box, the bootstrapping and key-switching keys replicated ONCE by an NCCL broadcast over NVLink, no collective per gate.
//
//   group_create(devices[], n)       one device context per GPU + one NCCL communicator per GPU (ncclCommInitAll, one process)
//   group_load_bk / _load_ksk        host -> device 0, ncclBroadcast (device 0 -> all), per-device transform into the NTT domain
//   group_keygen[_csprng]            both keys generated ON device 0, then the same broadcast
//   group_gate_batch[_async]         contiguous shards of ceil/floor(B / n) gates per GPU, every GPU fed from the caller's host
//                                    buffers on its own streams (tfhe_b200_gate_batch_async per context); group_sync waits
//   group_circuit_create / _run      a levelised netlist on all GPUs: wide levels sharded + outputs exchanged over NCCL, narrow
//                                    levels replicated (the one place of the path with a per-level exchange step)
// The per-device contexts stay reachable (group_ctx) for the device-pointer entry points.  One host thread at a time, like a ctx.
#include <cuda_runtime.h>
#include <nccl.h>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/tfhe_b200.h"

namespace {
constexpr size_t BK_WORDS = (size_t)635 * 12 * 1024;
constexpr size_t KSK_WORDS = (size_t)1024 * 8 * 3 * 636;
constexpr size_t CT_WORDS = 636;
thread_local std::string g_group_create_err;
}  // namespace

struct tfhe_b200_group {
    std::vector<int> devices;
    std::vector<tfhe_b200_ctx*> ctx;
    std::vector<ncclComm_t> comm;
    std::vector<cudaStream_t> stream;     // one per device: broadcasts and key loads
    std::vector<uint32_t*> stage;         // one per device: KSK_WORDS words, receive buffer of the broadcasts
    uint32_t mu = 0x20000000u;            // 1/8: the message of a trivial ciphertext (constant wires of a circuit)
    std::string err;
};

#define GCK(call)                                                                       \
    do {                                                                                \
        cudaError_t e_ = (call);                                                        \
        if (e_ != cudaSuccess) {                                                        \
            g->err = std::string(#call) + ": " + cudaGetErrorString(e_);                \
            return TFHE_B200_ERR_CUDA;                                                  \
        }                                                                               \
    } while (0)
#define GNCCL(call)                                                                     \
    do {                                                                                \
        ncclResult_t r_ = (call);                                                       \
        if (r_ != ncclSuccess) {                                                        \
            g->err = std::string(#call) + ": " + ncclGetErrorString(r_);                \
            return TFHE_B200_ERR_CUDA;                                                  \
        }                                                                               \
    } while (0)
#define GCTX(r, call)                                                                   \
    do {                                                                                \
        int rc_ = (call);                                                               \
        if (rc_ != TFHE_B200_OK) {                                                      \
            g->err = "device " + std::to_string(g->devices[r]) + ": " + tfhe_b200_last_error(g->ctx[r]); \
            return rc_;                                                                 \
        }                                                                               \
    } while (0)

// root's stage buffer holds `words` words: replicate them into every other device's stage buffer (one NCCL group call)
static int broadcast_stage(tfhe_b200_group* g, size_t words) {
    const int n = (int)g->ctx.size();
    if (n == 1) return TFHE_B200_OK;
    GNCCL(ncclGroupStart());
    for (int r = 0; r < n; r++) {
        ncclResult_t rr = ncclBroadcast(g->stage[r], g->stage[r], words, ncclUint32, 0, g->comm[r], g->stream[r]);
        if (rr != ncclSuccess) {
            ncclGroupEnd();
            g->err = std::string("ncclBroadcast: ") + ncclGetErrorString(rr);
            return TFHE_B200_ERR_CUDA;
        }
    }
    GNCCL(ncclGroupEnd());
    return TFHE_B200_OK;
}
static int sync_streams(tfhe_b200_group* g) {
    for (size_t r = 0; r < g->ctx.size(); r++) {
        GCK(cudaSetDevice(g->devices[r]));
        GCK(cudaStreamSynchronize(g->stream[r]));
    }
    return TFHE_B200_OK;
}
// stage buffers (already identical on every device) -> every context
static int load_from_stage(tfhe_b200_group* g, bool bk) {
    for (size_t r = 0; r < g->ctx.size(); r++) {
        GCK(cudaSetDevice(g->devices[r]));
        if (bk) GCTX(r, tfhe_b200_load_bk_device(g->ctx[r], g->stage[r], g->stream[r]));
        else GCTX(r, tfhe_b200_load_ksk_device(g->ctx[r], g->stage[r], g->stream[r]));
    }
    return sync_streams(g);
}

extern "C" {

const char* tfhe_b200_group_last_error(const tfhe_b200_group* g) { return g ? g->err.c_str() : g_group_create_err.c_str(); }

int tfhe_b200_group_create(const tfhe_b200_params* p, const int* devices, int ndev, tfhe_b200_group** out) {
    if (!out) { g_group_create_err = "group_create: out is NULL"; return TFHE_B200_ERR_PARAM; }
    *out = nullptr;
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have == 0) {
        g_group_create_err = "group_create: no CUDA device; there is no CPU fallback";
        return TFHE_B200_ERR_CUDA;
    }
    if (ndev <= 0 && !devices) ndev = have;   // all devices of the box
    if (ndev <= 0 || ndev > have) { g_group_create_err = "group_create: bad device count"; return TFHE_B200_ERR_PARAM; }
    tfhe_b200_group* g = new tfhe_b200_group();
    if (p) g->mu = p->mu;
    for (int r = 0; r < ndev; r++) {
        const int d = devices ? devices[r] : r;
        for (int q = 0; q < r; q++)
            if (g->devices[q] == d) { g_group_create_err = "group_create: duplicate device"; delete g; return TFHE_B200_ERR_PARAM; }
        g->devices.push_back(d);
    }
    auto bail = [&](int code, const std::string& msg) { g_group_create_err = msg; tfhe_b200_group_destroy(g); return code; };
    for (int r = 0; r < ndev; r++) {
        tfhe_b200_ctx* c = nullptr;
        const int rc = tfhe_b200_ctx_create(p, g->devices[r], &c);
        if (rc != TFHE_B200_OK) return bail(rc, std::string("group_create: device ") + std::to_string(g->devices[r]) + ": " + tfhe_b200_last_error(nullptr));
        g->ctx.push_back(c);
        cudaStream_t st = nullptr;
        uint32_t* buf = nullptr;
        if (cudaSetDevice(g->devices[r]) != cudaSuccess || cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess ||
            cudaMalloc(&buf, KSK_WORDS * 4) != cudaSuccess)
            return bail(TFHE_B200_ERR_CUDA, std::string("group_create: stream / staging buffer: ") + cudaGetErrorString(cudaGetLastError()));
        g->stream.push_back(st);
        g->stage.push_back(buf);
    }
    if (ndev > 1) {
        g->comm.resize(ndev);
        const ncclResult_t nr = ncclCommInitAll(g->comm.data(), ndev, g->devices.data());
        if (nr != ncclSuccess) { g->comm.clear(); return bail(TFHE_B200_ERR_CUDA, std::string("ncclCommInitAll: ") + ncclGetErrorString(nr)); }
    }
    *out = g;
    return TFHE_B200_OK;
}

int tfhe_b200_group_destroy(tfhe_b200_group* g) {
    if (!g) return TFHE_B200_ERR_PARAM;
    for (size_t r = 0; r < g->ctx.size(); r++) tfhe_b200_sync(g->ctx[r]);
    for (ncclComm_t c : g->comm) ncclCommDestroy(c);
    for (size_t r = 0; r < g->stage.size(); r++) {
        cudaSetDevice(g->devices[r]);
        cudaFree(g->stage[r]);
        cudaStreamDestroy(g->stream[r]);
    }
    for (tfhe_b200_ctx* c : g->ctx) tfhe_b200_ctx_destroy(c);
    delete g;
    return TFHE_B200_OK;
}

int tfhe_b200_group_size(const tfhe_b200_group* g) { return g ? (int)g->ctx.size() : 0; }
tfhe_b200_ctx* tfhe_b200_group_ctx(tfhe_b200_group* g, int rank) { return (g && rank >= 0 && rank < (int)g->ctx.size()) ? g->ctx[rank] : nullptr; }

int tfhe_b200_group_load_bk(tfhe_b200_group* g, const uint32_t* bk_host) {
    if (!g || !bk_host) return TFHE_B200_ERR_PARAM;
    GCK(cudaSetDevice(g->devices[0]));
    GCK(cudaMemcpyAsync(g->stage[0], bk_host, BK_WORDS * 4, cudaMemcpyHostToDevice, g->stream[0]));
    int rc = broadcast_stage(g, BK_WORDS);
    if (rc) return rc;
    return load_from_stage(g, true);
}
int tfhe_b200_group_load_ksk(tfhe_b200_group* g, const uint32_t* ksk_host) {
    if (!g || !ksk_host) return TFHE_B200_ERR_PARAM;
    GCK(cudaSetDevice(g->devices[0]));
    GCK(cudaMemcpyAsync(g->stage[0], ksk_host, KSK_WORDS * 4, cudaMemcpyHostToDevice, g->stream[0]));
    int rc = broadcast_stage(g, KSK_WORDS);
    if (rc) return rc;
    return load_from_stage(g, false);
}
// both keys generated on device 0, exported device-to-device into the staging buffer, broadcast, loaded everywhere else
static int replicate_from_root(tfhe_b200_group* g) {
    if (g->ctx.size() == 1) return TFHE_B200_OK;
    for (int pass = 0; pass < 2; pass++) {
        const bool bk = pass == 0;
        GCK(cudaSetDevice(g->devices[0]));
        if (bk) GCTX(0, tfhe_b200_export_bk_device(g->ctx[0], g->stage[0], g->stream[0]));
        else GCTX(0, tfhe_b200_export_ksk_device(g->ctx[0], g->stage[0], g->stream[0]));
        int rc = broadcast_stage(g, bk ? BK_WORDS : KSK_WORDS);
        if (rc) return rc;
        for (size_t r = 1; r < g->ctx.size(); r++) {
            GCK(cudaSetDevice(g->devices[r]));
            if (bk) GCTX(r, tfhe_b200_load_bk_device(g->ctx[r], g->stage[r], g->stream[r]));
            else GCTX(r, tfhe_b200_load_ksk_device(g->ctx[r], g->stage[r], g->stream[r]));
        }
        rc = sync_streams(g);
        if (rc) return rc;
    }
    return TFHE_B200_OK;
}
int tfhe_b200_group_keygen(tfhe_b200_group* g, uint64_t seed, const uint8_t* s0, const uint8_t* s1) {
    if (!g) return TFHE_B200_ERR_PARAM;
    GCTX(0, tfhe_b200_keygen_device(g->ctx[0], seed, s0, s1));
    return replicate_from_root(g);
}
int tfhe_b200_group_keygen_csprng(tfhe_b200_group* g, const uint8_t* key, const uint8_t* s0, const uint8_t* s1) {
    if (!g) return TFHE_B200_ERR_PARAM;
    GCTX(0, tfhe_b200_keygen_device_csprng(g->ctx[0], key, s0, s1));
    return replicate_from_root(g);
}

int tfhe_b200_group_reserve(tfhe_b200_group* g, size_t max_batch) {
    if (!g) return TFHE_B200_ERR_PARAM;
    const size_t n = g->ctx.size();
    for (size_t r = 0; r < n; r++) GCTX(r, tfhe_b200_reserve(g->ctx[r], (max_batch + n - 1) / n));
    return TFHE_B200_OK;
}
// gates [first, first + count) of a batch of B go to rank r: contiguous shards, the first B % n ranks take one more
void tfhe_b200_group_shard(const tfhe_b200_group* g, size_t B, int rank, size_t* first, size_t* count) {
    const size_t n = g ? g->ctx.size() : 1, base = B / n, rem = B % n, r = (size_t)rank;
    if (first) *first = r * base + (r < rem ? r : rem);
    if (count) *count = base + (r < rem ? 1 : 0);
}
int tfhe_b200_group_gate_batch_async(tfhe_b200_group* g, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B) {
    if (!g || !in0 || !out) return TFHE_B200_ERR_PARAM;
    for (size_t r = 0; r < g->ctx.size(); r++) {
        size_t first, count;
        tfhe_b200_group_shard(g, B, (int)r, &first, &count);
        if (count == 0) continue;
        GCTX(r, tfhe_b200_gate_batch_async(g->ctx[r], op, in0 + first * CT_WORDS, in1 ? in1 + first * CT_WORDS : nullptr, out + first * CT_WORDS, count));
    }
    return TFHE_B200_OK;
}
int tfhe_b200_group_sync(tfhe_b200_group* g) {
    if (!g) return TFHE_B200_ERR_PARAM;
    for (size_t r = 0; r < g->ctx.size(); r++) GCTX(r, tfhe_b200_sync(g->ctx[r]));
    return TFHE_B200_OK;
}
int tfhe_b200_group_gate_batch(tfhe_b200_group* g, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B) {
    const int rc = tfhe_b200_group_gate_batch_async(g, op, in0, in1, out, B);
    if (rc) { if (g) tfhe_b200_group_sync(g); return rc; }
    return tfhe_b200_group_sync(g);
}
// ---- circuits on a group: every device holds the whole wire table; a level of at least `shard_min` gates is cut
// into contiguous shards, one per device, whose outputs are exchanged before the next level -- each device broadcasts its
// shard of the level's output rows to the others (N ncclBroadcast in one NCCL group call = an all-gather with uneven shards,
// 2544 B per gate over NVLink) and every device scatters the rows into its wire table.  Narrower levels are evaluated by
// EVERY device on its own copy (the kernels are deterministic, the copies stay bit-identical): no exchange where a level
// fits one wave of the latency kernel anyway.  Everything is enqueued on one stream per device; the host waits once.
struct tfhe_b200_group_circuit {
    std::vector<tfhe_b200_circuit*> per_dev;
    std::vector<uint32_t*> wires, rows;
    std::vector<size_t> level_gates;
    size_t n_wires = 0, max_level = 0, shard_min = 0;
    uint64_t sharded_levels = 0, replicated_levels = 0;   // of the last run
    std::vector<uint32_t> trivial;                        // host source of the constant wires' ciphertexts
    int32_t* out_idx = nullptr;                           // device 0: the output wires of a run and their gathered rows
    uint32_t* out_rows = nullptr;
    size_t out_cap = 0;
};
// rows[k] = table[idx[k]] (one warp per 2544-byte ciphertext)
__global__ void gather_rows_kernel(const uint4* __restrict__ table, const int32_t* __restrict__ idx, uint4* __restrict__ rows, long n) {
    const long k = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= n) return;
    const uint4* src = table + (size_t)idx[k] * (CT_WORDS / 4);
    uint4* dst = rows + (size_t)k * (CT_WORDS / 4);
    for (int c = threadIdx.x & 31; c < (int)(CT_WORDS / 4); c += 32) dst[c] = src[c];
}
int tfhe_b200_group_circuit_destroy(tfhe_b200_group* g, tfhe_b200_group_circuit* c) {
    if (!g || !c) return TFHE_B200_ERR_PARAM;
    for (size_t r = 0; r < c->per_dev.size(); r++) {
        cudaSetDevice(g->devices[r]);
        if (c->per_dev[r]) tfhe_b200_circuit_destroy(g->ctx[r], c->per_dev[r]);
        if (r < c->wires.size()) cudaFree(c->wires[r]);
        if (r < c->rows.size()) cudaFree(c->rows[r]);
    }
    if (!g->devices.empty()) cudaSetDevice(g->devices[0]);
    cudaFree(c->out_idx);
    cudaFree(c->out_rows);
    delete c;
    return TFHE_B200_OK;
}
int tfhe_b200_group_circuit_create(tfhe_b200_group* g, size_t n_levels, const size_t* level_gates, const uint8_t* ops, const int32_t* in0,
                                   const int32_t* in1, const int32_t* out, size_t n_wires, size_t shard_min,
                                   tfhe_b200_group_circuit** circuit) {
    if (!g || !circuit) return TFHE_B200_ERR_PARAM;
    *circuit = nullptr;
    tfhe_b200_group_circuit* c = new tfhe_b200_group_circuit();
    c->n_wires = n_wires;
    if (shard_min == 0) {   // default: a level is cut up when it does not fit one wave of the 2-SM cluster latency kernel (#SMs / 2 gates)
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, g->devices[0]);
        shard_min = (size_t)sms / 2 + 1;
    }
    c->shard_min = shard_min;
    for (size_t l = 0; l < n_levels; l++) {
        c->level_gates.push_back(level_gates[l]);
        if (level_gates[l] > c->max_level) c->max_level = level_gates[l];
    }
    for (size_t r = 0; r < g->ctx.size(); r++) {
        tfhe_b200_circuit* pc = nullptr;
        const int rc = tfhe_b200_circuit_create(g->ctx[r], n_levels, level_gates, ops, in0, in1, out, n_wires, &pc);
        if (rc != TFHE_B200_OK) {
            g->err = "device " + std::to_string(g->devices[r]) + ": " + tfhe_b200_last_error(g->ctx[r]);
            tfhe_b200_group_circuit_destroy(g, c);
            return rc;
        }
        c->per_dev.push_back(pc);
        uint32_t *w = nullptr, *rows = nullptr;
        if (cudaSetDevice(g->devices[r]) != cudaSuccess || cudaMalloc(&w, (n_wires ? n_wires : 1) * CT_WORDS * 4) != cudaSuccess ||
            cudaMalloc(&rows, (c->max_level ? c->max_level : 1) * CT_WORDS * 4) != cudaSuccess) {
            g->err = std::string("group_circuit_create: ") + cudaGetErrorString(cudaGetLastError());
            cudaFree(w);
            tfhe_b200_group_circuit_destroy(g, c);
            return TFHE_B200_ERR_NOMEM;
        }
        c->wires.push_back(w);
        c->rows.push_back(rows);
    }
    *circuit = c;
    return TFHE_B200_OK;
}
// inputs_host [n_inputs][n+1] = wires 0 .. n_inputs-1 (uploaded to device 0 once and broadcast over NCCL); constant wires are
// set to trivial ciphertexts on every device; outputs_host [n_out][n+1] = the wires out_wires[], gathered on device 0.
static int group_circuit_run_impl(tfhe_b200_group* g, tfhe_b200_group_circuit* c, const uint32_t* inputs_host, size_t n_inputs,
                                  const int32_t* const_wires, const uint8_t* const_bits, size_t n_consts, const int32_t* out_wires, size_t n_out,
                                  uint32_t* outputs_host) {
    if (!g || !c || (n_inputs && !inputs_host) || (n_consts && (!const_wires || !const_bits)) || (n_out && (!out_wires || !outputs_host)))
        return TFHE_B200_ERR_PARAM;
    if (n_inputs > c->n_wires) { g->err = "group_circuit_run: more inputs than wires"; return TFHE_B200_ERR_PARAM; }
    for (size_t k = 0; k < n_consts; k++)
        if (const_wires[k] < 0 || (size_t)const_wires[k] >= c->n_wires) { g->err = "group_circuit_run: constant wire out of range"; return TFHE_B200_ERR_PARAM; }
    for (size_t k = 0; k < n_out; k++)
        if (out_wires[k] < 0 || (size_t)out_wires[k] >= c->n_wires) { g->err = "group_circuit_run: output wire out of range"; return TFHE_B200_ERR_PARAM; }
    const int n = (int)g->ctx.size();
    std::vector<uint32_t>& trivial = c->trivial;   // [0] = Zero, [1] = One: (b, a) = (-+1/8, 0)  (tlwe.rs:181-186); lives with the circuit:
    trivial.assign(2 * CT_WORDS, 0u);              // the asynchronous copies below may still read it when an error returns early
    trivial[0] = 0u - g->mu;
    trivial[CT_WORDS] = g->mu;
    for (int r = 0; r < n; r++) {
        GCK(cudaSetDevice(g->devices[r]));
        GCK(cudaMemsetAsync(c->wires[r], 0, c->n_wires * CT_WORDS * 4, g->stream[r]));
    }
    if (n_inputs) {
        GCK(cudaSetDevice(g->devices[0]));
        GCK(cudaMemcpyAsync(c->wires[0], inputs_host, n_inputs * CT_WORDS * 4, cudaMemcpyHostToDevice, g->stream[0]));
        if (n > 1) {
            GNCCL(ncclGroupStart());
            for (int r = 0; r < n; r++) {
                const ncclResult_t rr = ncclBroadcast(c->wires[r], c->wires[r], n_inputs * CT_WORDS, ncclUint32, 0, g->comm[r], g->stream[r]);
                if (rr != ncclSuccess) { ncclGroupEnd(); g->err = std::string("ncclBroadcast (inputs): ") + ncclGetErrorString(rr); return TFHE_B200_ERR_CUDA; }
            }
            GNCCL(ncclGroupEnd());
        }
    }
    for (int r = 0; r < n; r++) {
        GCK(cudaSetDevice(g->devices[r]));
        for (size_t k = 0; k < n_consts; k++)
            GCK(cudaMemcpyAsync(c->wires[r] + (size_t)const_wires[k] * CT_WORDS, trivial.data() + (const_bits[k] ? CT_WORDS : 0), CT_WORDS * 4,
                                cudaMemcpyHostToDevice, g->stream[r]));
    }
    c->sharded_levels = c->replicated_levels = 0;
    for (size_t l = 0; l < c->level_gates.size(); l++) {
        const size_t width = c->level_gates[l];
        if (width == 0) continue;
        if (n == 1 || width < c->shard_min) {
            for (int r = 0; r < n; r++)
                GCTX(r, tfhe_b200_circuit_run_level_device(g->ctx[r], c->per_dev[r], l, 0, width, c->wires[r], nullptr, g->stream[r]));
            c->replicated_levels++;
            continue;
        }
        for (int r = 0; r < n; r++) {
            size_t first, count;
            tfhe_b200_group_shard(g, width, r, &first, &count);
            GCTX(r, tfhe_b200_circuit_run_level_device(g->ctx[r], c->per_dev[r], l, first, count, c->wires[r], c->rows[r] + first * CT_WORDS,
                                                      g->stream[r]));
        }
        GNCCL(ncclGroupStart());
        for (int root = 0; root < n; root++) {
            size_t first, count;
            tfhe_b200_group_shard(g, width, root, &first, &count);
            if (count == 0) continue;
            for (int r = 0; r < n; r++) {
                uint32_t* p = c->rows[r] + first * CT_WORDS;
                const ncclResult_t rr = ncclBroadcast(p, p, count * CT_WORDS, ncclUint32, root, g->comm[r], g->stream[r]);
                if (rr != ncclSuccess) {
                    ncclGroupEnd();
                    g->err = std::string("ncclBroadcast (level outputs): ") + ncclGetErrorString(rr);
                    return TFHE_B200_ERR_CUDA;
                }
            }
        }
        GNCCL(ncclGroupEnd());
        for (int r = 0; r < n; r++)
            GCTX(r, tfhe_b200_circuit_scatter_level_device(g->ctx[r], c->per_dev[r], l, c->rows[r], c->wires[r], g->stream[r]));
        c->sharded_levels++;
    }
    if (n_out) {
        GCK(cudaSetDevice(g->devices[0]));
        if (n_out > c->out_cap) {
            GCK(cudaStreamSynchronize(g->stream[0]));
            cudaFree(c->out_idx); cudaFree(c->out_rows);
            c->out_idx = nullptr; c->out_rows = nullptr; c->out_cap = 0;
            GCK(cudaMalloc(&c->out_idx, n_out * sizeof(int32_t)));
            GCK(cudaMalloc(&c->out_rows, n_out * CT_WORDS * 4));
            c->out_cap = n_out;
        }
        GCK(cudaMemcpyAsync(c->out_idx, out_wires, n_out * sizeof(int32_t), cudaMemcpyHostToDevice, g->stream[0]));
        gather_rows_kernel<<<(unsigned)((n_out + 7) / 8), 256, 0, g->stream[0]>>>(reinterpret_cast<const uint4*>(c->wires[0]), c->out_idx,
                                                                                 reinterpret_cast<uint4*>(c->out_rows), (long)n_out);
        GCK(cudaGetLastError());
        GCK(cudaMemcpyAsync(outputs_host, c->out_rows, n_out * CT_WORDS * 4, cudaMemcpyDeviceToHost, g->stream[0]));
    }
    return sync_streams(g);
}
int tfhe_b200_group_circuit_run(tfhe_b200_group* g, tfhe_b200_group_circuit* c, const uint32_t* inputs_host, size_t n_inputs,
                                const int32_t* const_wires, const uint8_t* const_bits, size_t n_consts, const int32_t* out_wires, size_t n_out,
                                uint32_t* outputs_host) {
    const int rc = group_circuit_run_impl(g, c, inputs_host, n_inputs, const_wires, const_bits, n_consts, out_wires, n_out, outputs_host);
    if (rc != TFHE_B200_OK && g) {   // an error half way: nothing enqueued so far may outlive the caller's buffers
        const std::string first = g->err;
        for (size_t r = 0; r < g->ctx.size(); r++) {
            cudaSetDevice(g->devices[r]);
            cudaStreamSynchronize(g->stream[r]);
        }
        g->err = first;
    }
    return rc;
}
int tfhe_b200_group_circuit_stats(const tfhe_b200_group_circuit* c, uint64_t* sharded_levels, uint64_t* replicated_levels, size_t* shard_min) {
    if (!c) return TFHE_B200_ERR_PARAM;
    if (sharded_levels) *sharded_levels = c->sharded_levels;
    if (replicated_levels) *replicated_levels = c->replicated_levels;
    if (shard_min) *shard_min = c->shard_min;
    return TFHE_B200_OK;
}
// pinned host memory that every device of the box can copy from / to asynchronously (pageable buffers serialise the shards)
int tfhe_b200_host_alloc(void** out, size_t bytes) {
    if (!out) return TFHE_B200_ERR_PARAM;
    return cudaHostAlloc(out, bytes, cudaHostAllocPortable) == cudaSuccess ? TFHE_B200_OK : TFHE_B200_ERR_NOMEM;
}
int tfhe_b200_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? TFHE_B200_OK : TFHE_B200_ERR_CUDA; }

}  // extern "C"
