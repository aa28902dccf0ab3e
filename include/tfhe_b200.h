/*
 * tfhe_b200.h -- C ABI of the B200-native TFHE gate-bootstrapping engine (librustfhe_b200.so).
 *
 * Drop-in boundary for the bootstrapped-HomNAND path of hideki1217/rusTfhe.  In the reference the only
 * language boundary is the FFT binding
 *     utils/src/spqlios.rs:18-32  <->  utils/src/spqlios/spqlios-wrapper.cpp:4-53   (8 functions, opaque handle)
 * crossed 5080 times per gate.  Here the boundary moves up: ONE call submits a BATCH of gates and everything
 * under hom_nand::tfhe::TFHE::{hom_nand,hom_and,hom_or,hom_xor,hom_not,hom_mux} (hom_nand/src/tfhe.rs:27-113) runs
 * in CUDA on sm_100a.  Same style as the reference binding: plain pointers and sizes, opaque handle, caller
 * owns every buffer.  Differences, on purpose: every function returns an int status (the reference returns void
 * and abort()s, spqlios-fft-impl.cpp:92-97) and destroy really frees (the reference leaks,
 * spqlios-wrapper.cpp:14-16).  There is NO CPU fallback: every compute entry fails with TFHE_B200_ERR_CUDA when no
 * sm_100 device is usable.
 *
 * Flat little-endian u32 layouts (Torus32 = u32, wrapping arithmetic, utils/src/math.rs:489-539):
 *   TLWE lv0 ciphertext : [n+1]      word 0 = b ("cipher"), words 1..n = a ("p_key")       hom_nand/src/tlwe.rs:19-22
 *   TLWE lv1 ciphertext : [N+1]      same, dimension N                                      hom_nand/src/trlwe.rs:110-121
 *   TRLWE               : [2][N]     poly 0 = b ("cipher"), poly 1 = a ("p_key")            hom_nand/src/trlwe.rs:13-16
 *   TRGSW               : [2l][2][N] rows 0..l carry mu/Bg^(j+1) on b, rows l..2l on a      hom_nand/src/trgsw.rs:23-26,213-229
 *   bootstrapping key   : [n][2l][2][N]   BK_i = TRGSW_{s1}(s0_i), TORUS domain            hom_nand/src/tfhe.rs:116-126
 *   key-switching key   : [N][t][3][n+1]  entry (i,l,d-1) = TLWE_{s0}(d*s1_i / 2^(2(l+1))), d=1..3
 *                                         (the reference also stores an unreachable d=4)   hom_nand/src/tlwe.rs:243-283
 * Default parameters: n=635, N=1024, l=3, Bgbit=6, t=8, basebit=2, mu=1/8 (tlwe.rs:175-186, trlwe.rs:76,
 * trgsw.rs:112-115, tfhe.rs:16-17).  This build is specialised for exactly these (other values -> ERR_PARAM).
 *
 * Threading: a ctx may be used by one host thread at a time (the reference handle is not thread safe either,
 * spqlios-fft.h:25-30).  "_device" variants take device pointers and a cudaStream_t (as void*), enqueue
 * asynchronously and never synchronise; calls issued on different streams may overlap on the device (workspaces are
 * handed out from an event-guarded ring).  The host-pointer variants copy in, run, copy out and synchronise.
 */
#ifndef TFHE_B200_H
#define TFHE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFHE_B200_OK 0
#define TFHE_B200_ERR_PARAM 1    /* bad argument / unsupported parameter set */
#define TFHE_B200_ERR_CUDA 2     /* CUDA runtime error (see tfhe_b200_last_error) */
#define TFHE_B200_ERR_STATE 3    /* keys not loaded */
#define TFHE_B200_ERR_NOMEM 4
#define TFHE_B200_ERR_IO 5       /* file format / I/O error (see tfhe_b200_file_last_error) */

/* gate opcodes: the linear pre-combination applied before the bootstrap (hom_nand/src/tfhe.rs:27-71) */
#define TFHE_B200_NAND 0   /* (1/8,0) - (c0+c1)          tfhe.rs:41-47 */
#define TFHE_B200_AND 1    /* (c0+c1) - (1/8,0)          tfhe.rs:48-54 */
#define TFHE_B200_OR 2     /* (c0+c1) + (1/8,0)          tfhe.rs:55-61 */
#define TFHE_B200_XOR 3    /* 2(c0+c1) + (1/4,0)         tfhe.rs:62-68 */
#define TFHE_B200_NOT 4    /* -c0 (still bootstrapped)   tfhe.rs:69-71 */
#define TFHE_B200_COPY 5   /* bootstrap(c0)              tfhe.rs:73-80 */
#define TFHE_B200_ANDNY 6  /* hom_and(-c0, c1)           tfhe.rs:34    */

#define TFHE_B200_MASK_FAITHFUL 0x02084000u /* Torus32::make_decomp_mask(3,6) as it evaluates, math.rs:542-560 */
#define TFHE_B200_MASK_CORRECTED 0x02082000u /* the mask the reference's decomposition KATs pin, math.rs:582-591 */

typedef struct tfhe_b200_ctx tfhe_b200_ctx;

typedef struct {
    int32_t n, N, l, bgbit, ks_t, ks_basebit;
    uint32_t mu;          /* test-vector coefficient / message amplitude, 0x20000000 = 1/8 */
    uint32_t decomp_mask; /* TFHE_B200_MASK_FAITHFUL by default */
} tfhe_b200_params;

typedef struct {
    uint64_t kernel_launches;  /* kernels of this library launched since ctx creation */
    float last_blind_rotate_ms; /* device time of the most recent blind-rotate kernel (CUDA events on its stream) */
    float last_keyswitch_ms;
    float avg_blind_rotate_ms;  /* mean over the last timed_launches (<= 64) gate batches since reset_stats */
    float avg_keyswitch_ms;
    uint64_t timed_launches;
    uint64_t last_batch;
    int32_t gates_per_cta, sm_count;
    uint64_t device_key_bytes;
    int32_t key_slices;         /* arithmetic mode: 1 = FFT64, 2 or 3 = NTT key slices (tfhe_b200_set_key_slices) */
    int32_t reserved;
} tfhe_b200_stats;

/* ---- lifecycle (replaces Spqlios_new / Spqlios_destructor, spqlios-wrapper.cpp:9-16) ---- */
int tfhe_b200_default_params(tfhe_b200_params* p);
int tfhe_b200_ctx_create(const tfhe_b200_params* p /* NULL = defaults */, int device, tfhe_b200_ctx** out);
int tfhe_b200_ctx_destroy(tfhe_b200_ctx* ctx);
const char* tfhe_b200_last_error(const tfhe_b200_ctx* ctx /* NULL = last error of a failed ctx_create */);
int tfhe_b200_set_decomp_mask(tfhe_b200_ctx* ctx, uint32_t mask);
/* Arithmetic of the polynomial products of the gate path (every mode returns the same ciphertext bits for honestly generated keys;
 * the parity tests compare all three against the exact-integer oracle):
 *   1 (default) FFT64: one f64 complex transform of the folded polynomial, 6 + 2 transforms per CMUX, every product rounded to
 *     the EXACT integer (the reference computes the same products with an f64 FFT, fft_processor_spqlios.cpp:58-183).  Measured
 *     rounding margin on uniform keys: |value - nearest integer| < 2^-8 against the 1/2 that would flip a bit (DESIGN.md
 *     section 2; tests/test_host_logic.py asserts it).  Used by every gate batch, whatever its size; the step-level entry
 *     points (external product, cmux) run mode 2.
 *   2 NTT over a 29-bit prime, two 16-bit key slices, 6 + 4 transforms per CMUX; exact unless a slice sum exceeds 9.8 standard
 *     deviations (about 1e-22 per coefficient, 3e-16 per gate over the key's masks).
 *   3 NTT, three 11-bit key slices, 6 + 6 transforms: exact in the worst case (any key, any digits).
 * Switching re-transforms the loaded bootstrapping key. */
int tfhe_b200_set_key_slices(tfhe_b200_ctx* ctx, int slices);
/* How a full batch (more than two gates per SM) is cut into CTAs.  AUTO (default): decided per call -- if an earlier batch is
 * still running on another stream the batch is cut into 4-gate CTAs only (the next batch back-fills the last wave), otherwise
 * it is dealt evenly over whole waves (best for a batch that has the device to itself).  STREAMED / LONE force either; a
 * caller that keeps several batches in flight sets STREAMED so that the first batch of a burst is cut the same way. */
#define TFHE_B200_OVERLAP_AUTO (-1)
#define TFHE_B200_OVERLAP_LONE 0
#define TFHE_B200_OVERLAP_STREAMED 1
int tfhe_b200_set_batch_overlap(tfhe_b200_ctx* ctx, int mode);
int tfhe_b200_get_stats(tfhe_b200_ctx* ctx, tfhe_b200_stats* out); /* waits for the recorded events */
int tfhe_b200_reset_stats(tfhe_b200_ctx* ctx);

/* ---- keys: BootstrappingKey::new / KeySwitchingKey::new products (tfhe.rs:119-126, tlwe.rs:247-277) ----
 * load_bk transforms the torus-domain key on the device into the NTT domain (replaces TRGSWRepF::from,
 * trgsw.rs:68-76).  The *_device variants read device memory (e.g. the destination of an NCCL broadcast) and are
 * asynchronous on the caller's stream; batches issued afterwards wait for the load, whatever stream they run on.  Loading
 * keys while batches are still in flight is the caller's race: tfhe_b200_sync first. */
int tfhe_b200_load_bk(tfhe_b200_ctx* ctx, const uint32_t* bk_host);
int tfhe_b200_load_bk_device(tfhe_b200_ctx* ctx, const uint32_t* bk_dev, void* stream);
int tfhe_b200_load_ksk(tfhe_b200_ctx* ctx, const uint32_t* ksk_host);
int tfhe_b200_load_ksk_device(tfhe_b200_ctx* ctx, const uint32_t* ksk_dev, void* stream);

/* ---- the hot path: batched bootstrapped gates (TFHE::hom_* , tfhe.rs:27-80) ----
 * in0,in1,out : [B][n+1].  in1 is ignored (may be NULL) for NOT/COPY.  out may alias an input. */
int tfhe_b200_gate_batch(tfhe_b200_ctx* ctx, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B);
int tfhe_b200_gate_batch_device(tfhe_b200_ctx* ctx, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out,
                                size_t B, void* stream);
/* Asynchronous host-pointer form: H2D copy, gate batch and D2H copy are enqueued on one of the context's internal
 * streams and the call returns; buffers must stay valid (and should be pinned) until tfhe_b200_sync returns.
 * Consecutive async batches overlap on the device (the tail of one blind rotation runs under the head of the next). */
int tfhe_b200_gate_batch_async(tfhe_b200_ctx* ctx, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B);
int tfhe_b200_sync(tfhe_b200_ctx* ctx); /* waits for every batch this context has enqueued (any stream) */
/* optional: pre-allocate all internal workspaces for batches of up to max_batch gates (otherwise they grow on demand,
 * and cudaMalloc blocks while earlier batches are still running) */
int tfhe_b200_reserve(tfhe_b200_ctx* ctx, size_t max_batch);
/* One launch for a batch of gates with DIFFERENT opcodes -- one level of a levelised circuit (the reference evaluates
 * nander expressions one gate at a time, nander/src/lib.rs:72-89).  ops[g] is one of the opcodes above; in1[g] is ignored
 * for NOT / COPY gates. */
int tfhe_b200_gate_batch_mixed(tfhe_b200_ctx* ctx, const uint8_t* ops /*[B]*/, const uint32_t* in0, const uint32_t* in1,
                               uint32_t* out, size_t B);
int tfhe_b200_gate_batch_mixed_device(tfhe_b200_ctx* ctx, const uint8_t* ops_dev, const uint32_t* in0, const uint32_t* in1,
                                      uint32_t* out, size_t B, void* stream);
/* Device-resident circuits: a levelised gate netlist (wire indices into one table of ciphertexts) uploaded once; a run
 * enqueues one blind-rotate + one key-switch launch per level on `stream` and never returns to the host in between.
 * Arrays are the levels concatenated: ops[g], in0[g], in1[g] (ignored for NOT / COPY), out[g] are wire indices < n_wires;
 * a gate may only read wires written by earlier levels (or inputs / constants). */
typedef struct tfhe_b200_circuit tfhe_b200_circuit;
int tfhe_b200_circuit_create(tfhe_b200_ctx* ctx, size_t n_levels, const size_t* level_gates, const uint8_t* ops,
                             const int32_t* in0, const int32_t* in1, const int32_t* out, size_t n_wires,
                             tfhe_b200_circuit** circuit);
int tfhe_b200_circuit_run_device(tfhe_b200_ctx* ctx, const tfhe_b200_circuit* circuit, uint32_t* wires_dev /*[n_wires][n+1]*/,
                                 void* stream);
int tfhe_b200_circuit_destroy(tfhe_b200_ctx* ctx, tfhe_b200_circuit* circuit);
/* Pieces of a run, for evaluators that spread one level over several devices (the group entry points below use them):
 * the shape of a circuit; gates [first, first + count) of one level, results either to their wires (rows_out == NULL) or
 * to row k = gate first + k of rows_out with the wire table only read; and the scatter of a whole level's rows
 * (rows_dev[k] = output of gate k of the level) into the wire table. */
int tfhe_b200_circuit_shape(const tfhe_b200_circuit* circuit, size_t* n_levels, size_t* n_wires, size_t* max_level_gates);
int tfhe_b200_circuit_level_gates(const tfhe_b200_circuit* circuit, size_t level, size_t* gates);
int tfhe_b200_circuit_run_level_device(tfhe_b200_ctx* ctx, const tfhe_b200_circuit* circuit, size_t level, size_t first, size_t count,
                                       uint32_t* wires_dev, uint32_t* rows_out /* NULL or [count][n+1] */, void* stream);
int tfhe_b200_circuit_scatter_level_device(tfhe_b200_ctx* ctx, const tfhe_b200_circuit* circuit, size_t level,
                                           const uint32_t* rows_dev /*[level gates][n+1]*/, uint32_t* wires_dev, void* stream);
int tfhe_b200_bootstrap_batch(tfhe_b200_ctx* ctx, const uint32_t* in, uint32_t* out, size_t B); /* TFHE::bootstrap */
/* hom_mux(control, input_0, input_1) = (input_1 & control) | (input_0 & !control): three bootstraps, tfhe.rs:27-40 */
int tfhe_b200_mux_batch(tfhe_b200_ctx* ctx, const uint32_t* control, const uint32_t* in0, const uint32_t* in1,
                        uint32_t* out, size_t B);
int tfhe_b200_mux_batch_device(tfhe_b200_ctx* ctx, const uint32_t* control, const uint32_t* in0, const uint32_t* in1,
                               uint32_t* out, size_t B, void* stream);

/* ---- step-level entries (parity ladder, micro-benchmarks; host pointers) ---- */
/* TFHE::blind_rotate restricted to the first nsteps key elements (nsteps = n for the full rotation): [B][2][N] */
int tfhe_b200_blind_rotate_batch(tfhe_b200_ctx* ctx, const uint32_t* in /*[B][n+1]*/, int nsteps, uint32_t* out_trlwe,
                                 size_t B);
/* blind_rotate + sample_extract_index(0): [B][N+1]  (tfhe.rs:81-88) */
int tfhe_b200_bootstrap_lv1_batch(tfhe_b200_ctx* ctx, const uint32_t* in, uint32_t* out_lwe1, size_t B);
/* TLWERep::identity_key_switch: [B][N+1] -> [B][n+1]  (tlwe.rs:43-73) */
int tfhe_b200_keyswitch_batch(tfhe_b200_ctx* ctx, const uint32_t* lwe1, uint32_t* out, size_t B);
/* TRGSWRepF::cross: out[g] = trgsw[g % ntrgsw] (x) trlwe[g]   (trgsw.rs:264-306); trgsw in the torus domain */
int tfhe_b200_external_product_batch(tfhe_b200_ctx* ctx, const uint32_t* trgsw /*[ntrgsw][2l][2][N]*/, size_t ntrgsw,
                                     const uint32_t* trlwe /*[B][2][N]*/, uint32_t* out /*[B][2][N]*/, size_t B);
/* Polynomial::fft_cross / Spqlios_poly_mul replacement (math.rs:337-347, spqlios-wrapper.cpp:38-53), exact:
 * out[g] = a[g] * d[g] mod (X^N+1, 2^32), a torus, |d| <= 192 */
int tfhe_b200_negacyclic_mul_batch(tfhe_b200_ctx* ctx, const uint32_t* a /*[B][N]*/, const int32_t* d /*[B][N]*/,
                                   uint32_t* out /*[B][N]*/, size_t B);

/* device-pointer forms of the two micro-benchmark entries (enqueue on `stream`, no synchronisation) */
int tfhe_b200_external_product_batch_device(tfhe_b200_ctx* ctx, const uint32_t* trgsw_dev, size_t ntrgsw,
                                            const uint32_t* trlwe_dev, uint32_t* out_dev, size_t B, void* stream);
int tfhe_b200_negacyclic_mul_batch_device(tfhe_b200_ctx* ctx, const uint32_t* a_dev, const int32_t* d_dev, uint32_t* out_dev,
                                          size_t B, void* stream);

/* ---- host-side key generation / encryption (TFHE::new, Cryptor::{encrypto,decrypto}; tfhe.rs:21-25,
 * tlwe.rs:197-241,247-277, trgsw.rs:117-139,213-229, trlwe.rs:127-137, digest.rs:14-33).
 *
 * PRODUCTION entry points (*_csprng): ChaCha20 keyed with 256 bits -- the role of the reference's rand::thread_rng, an
 * OS-seeded ChaCha generator (math.rs:421-476).  key = 32 bytes, or NULL: the call draws a fresh key from getrandom(2), so no
 * (key, index) pair is ever reused.  Callers that pass their own key must never use one key for two calls of the same
 * function, and must not derive the secret-key generator key and the public-material keys from one another. */
int tfhe_b200_random_bytes(uint8_t* out, size_t len);   /* getrandom(2) */
int tfhe_b200_keygen_secret_csprng(const uint8_t* key /*[32] or NULL*/, uint8_t* s0 /*[n]*/, uint8_t* s1 /*[N]*/);
int tfhe_b200_keygen_bk_csprng(const uint8_t* key /*[32] or NULL*/, const uint8_t* s0, const uint8_t* s1, uint32_t* bk);
int tfhe_b200_keygen_ksk_csprng(const uint8_t* key /*[32] or NULL*/, const uint8_t* s0, const uint8_t* s1, uint32_t* ksk);
int tfhe_b200_encrypt_bits_csprng(const uint8_t* key /*[32] or NULL*/, const uint8_t* s0, const uint8_t* bits, size_t B,
                                  uint32_t* out /*[B][n+1]*/);
/* DETERMINISTIC TEST entry points: SplitMix64 of a 64-bit seed.  NOT SECURE -- two public mask words reveal the generator
 * state, and with it every noise term and both secret keys; equal (seed, index) pairs repeat masks and noise.  They exist
 * only so that parity tests can give the oracle, the host and the device identical key material (the reference cannot be
 * seeded at all).  Never use them for data that needs protecting. */
int tfhe_b200_keygen_secret(uint64_t seed, uint8_t* s0 /*[n]*/, uint8_t* s1 /*[N]*/);
int tfhe_b200_keygen_bk(uint64_t seed, const uint8_t* s0, const uint8_t* s1, uint32_t* bk);
int tfhe_b200_keygen_ksk(uint64_t seed, const uint8_t* s0, const uint8_t* s1, uint32_t* ksk);
int tfhe_b200_encrypt_bits(uint64_t seed, uint64_t ct_index0, const uint8_t* s0, const uint8_t* bits, size_t B,
                           uint32_t* out /*[B][n+1]*/);
int tfhe_b200_phase(const uint8_t* s0, const uint32_t* ct, size_t B, uint32_t* phase);
int tfhe_b200_decrypt_bits(const uint8_t* s0, const uint32_t* ct, size_t B, uint8_t* bits);

/* ---- device-side key generation / encryption / decryption (same seeded generator as the host functions above: the
 * results are bit-identical to tfhe_b200_keygen_bk / _keygen_ksk / _encrypt_bits with the same seed).  Replaces the
 * host loops of BootstrappingKey::new (3810 TRLWE encryptions, tfhe.rs:119-126, trgsw.rs:117-139) and
 * KeySwitchingKey::new (24576 TLWE encryptions, tlwe.rs:247-277) that the reference marks "TODO: parallelise". ---- */
int tfhe_b200_keygen_device_csprng(tfhe_b200_ctx* ctx, const uint8_t* key /*[32] or NULL = getrandom*/, const uint8_t* s0 /*[n] host*/,
                                   const uint8_t* s1 /*[N] host*/);
int tfhe_b200_encrypt_bits_device_csprng(tfhe_b200_ctx* ctx, const uint8_t* key /*[32] or NULL = getrandom*/, const uint8_t* s0 /*host*/,
                                         const uint8_t* bits_dev /*[B] device*/, size_t B, uint32_t* out_dev /*[B][n+1] device*/,
                                         void* stream);
/* deterministic test forms (INSECURE, see above) */
int tfhe_b200_keygen_device(tfhe_b200_ctx* ctx, uint64_t seed, const uint8_t* s0 /*[n] host*/, const uint8_t* s1 /*[N] host*/);
int tfhe_b200_export_bk(tfhe_b200_ctx* ctx, uint32_t* bk_host /*[n][2l][2][N] torus domain*/);
int tfhe_b200_export_ksk(tfhe_b200_ctx* ctx, uint32_t* ksk_host /*[N][t][3][n+1]*/);
/* device-to-device forms (e.g. into the send buffer of the NCCL broadcast that replicates the keys) */
int tfhe_b200_export_bk_device(tfhe_b200_ctx* ctx, uint32_t* bk_dev, void* stream);
int tfhe_b200_export_ksk_device(tfhe_b200_ctx* ctx, uint32_t* ksk_dev, void* stream);
int tfhe_b200_encrypt_bits_device(tfhe_b200_ctx* ctx, uint64_t seed, uint64_t ct_index0, const uint8_t* s0 /*host*/,
                                  const uint8_t* bits_dev /*[B] device*/, size_t B, uint32_t* out_dev /*[B][n+1] device*/,
                                  void* stream);
int tfhe_b200_decrypt_bits_device(tfhe_b200_ctx* ctx, const uint8_t* s0 /*host*/, const uint32_t* ct_dev, size_t B,
                                  uint8_t* bits_dev /*[B] or NULL*/, uint32_t* phase_dev /*[B] or NULL*/, void* stream);

/* ---- remaining scheme surface below the gate level (host pointers, synchronous) ---- */
/* TRGSWRep::cmux(rep_1, rep_0) = cross(rep_1 - rep_0) + rep_0  (trgsw.rs:315-330); out[g] uses trgsw[g % ntrgsw] */
int tfhe_b200_cmux_batch(tfhe_b200_ctx* ctx, const uint32_t* trgsw /*[ntrgsw][2l][2][N]*/, size_t ntrgsw,
                         const uint32_t* rep1 /*[B][2][N]*/, const uint32_t* rep0 /*[B][2][N]*/, uint32_t* out, size_t B);
/* TRLWERep::sample_extract_index(index) (trlwe.rs:110-121): [B][2][N] -> [B][N+1] */
int tfhe_b200_sample_extract_batch(tfhe_b200_ctx* ctx, const uint32_t* trlwe, int index, uint32_t* out_lwe1, size_t B);

/* ---- flat little-endian file format for keys and ciphertexts (the reference has no serialisation; layout in
 * rustfhe_b200/csrc/wire.cpp: 64-byte header + the C-ABI layouts above + FNV-1a checksum) ---- */
#define TFHE_B200_FILE_SECRET 1 /* s0[n] bytes then s1[N] bytes */
#define TFHE_B200_FILE_BK 2     /* torus-domain bootstrapping key */
#define TFHE_B200_FILE_KSK 3
#define TFHE_B200_FILE_TLWE0 4  /* count x [n+1] */
#define TFHE_B200_FILE_TLWE1 5  /* count x [N+1] */
#define TFHE_B200_FILE_TRLWE 6  /* count x [2][N] */
#define TFHE_B200_FILE_TRGSW 7  /* count x [2l][2][N] */
int tfhe_b200_file_write(const char* path, int kind, const void* payload, uint64_t count);
int tfhe_b200_file_info(const char* path, int* kind, uint64_t* count, uint64_t* payload_bytes);
int tfhe_b200_file_read(const char* path, int kind, void* payload, uint64_t payload_bytes);
const char* tfhe_b200_file_last_error(void);

/* ---- device groups: ONE process drives several B200s of a box (rustfhe_b200/csrc/group.cu).  The reference has no
 * counterpart (TFHE::new keeps one key pair in host memory, tfhe.rs:21-25, everything runs on the calling thread); this is
 * independent gates of a batch are sharded over the GPUs (contiguous shards, the first B % n devices take
 * one gate more), the bootstrapping and key-switching keys are replicated ONCE by an ncclBroadcast over NVLink (device 0 ->
 * all, then every device transforms the bootstrapping key into its NTT domain locally); no collective runs per gate.
 * devices == NULL and ndev <= 0: every device of the box.  One host thread at a time per group. ---- */
typedef struct tfhe_b200_group tfhe_b200_group;
int tfhe_b200_group_create(const tfhe_b200_params* p /* NULL = defaults */, const int* devices, int ndev, tfhe_b200_group** out);
int tfhe_b200_group_destroy(tfhe_b200_group* g);
const char* tfhe_b200_group_last_error(const tfhe_b200_group* g /* NULL = last error of a failed group_create */);
int tfhe_b200_group_size(const tfhe_b200_group* g);
tfhe_b200_ctx* tfhe_b200_group_ctx(tfhe_b200_group* g, int rank);   /* the per-device context, for the device-pointer entry points */
int tfhe_b200_group_load_bk(tfhe_b200_group* g, const uint32_t* bk_host /*[n][2l][2][N]*/);
int tfhe_b200_group_load_ksk(tfhe_b200_group* g, const uint32_t* ksk_host /*[N][t][3][n+1]*/);
/* both keys generated on device 0 (tfhe_b200_keygen_device[_csprng]), then broadcast */
int tfhe_b200_group_keygen_csprng(tfhe_b200_group* g, const uint8_t* key /*[32] or NULL = getrandom*/, const uint8_t* s0, const uint8_t* s1);
int tfhe_b200_group_keygen(tfhe_b200_group* g, uint64_t seed /* deterministic TEST generator, INSECURE */, const uint8_t* s0, const uint8_t* s1);
int tfhe_b200_group_reserve(tfhe_b200_group* g, size_t max_batch /* whole batch, all devices */);
void tfhe_b200_group_shard(const tfhe_b200_group* g, size_t B, int rank, size_t* first, size_t* count);
/* host pointers; in1 may be NULL for NOT / COPY.  _async returns once every shard is enqueued on its device (buffers should be
 * pinned -- tfhe_b200_host_alloc -- or the host-to-device copies of the shards serialise); tfhe_b200_group_sync waits. */
int tfhe_b200_group_gate_batch(tfhe_b200_group* g, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B);
int tfhe_b200_group_gate_batch_async(tfhe_b200_group* g, int op, const uint32_t* in0, const uint32_t* in1, uint32_t* out, size_t B);
int tfhe_b200_group_sync(tfhe_b200_group* g);
/* A levelised netlist (arguments as tfhe_b200_circuit_create) on every device of the group.  A level of at least shard_min
 * gates (0 = default: more than one wave of the latency kernel, #SMs / 2 + 1) is cut into contiguous shards, one per device,
 * and its output ciphertexts are exchanged over NCCL before the next level (2544 B per gate); narrower levels are evaluated
 * by every device on its own copy of the wire table, without an exchange.  The reference evaluates a logic expression
 * depth-first, one gate at a time, on the calling thread (nander/src/lib.rs:72-89).
 * _run: inputs_host [n_inputs][n+1] are the wires 0 .. n_inputs-1 (one host-to-device copy, then an NCCL broadcast); the wires
 * const_wires[k] are set to the trivial ciphertext of const_bits[k]; outputs_host [n_out][n+1] receives the wires out_wires[]. */
typedef struct tfhe_b200_group_circuit tfhe_b200_group_circuit;
int tfhe_b200_group_circuit_create(tfhe_b200_group* g, size_t n_levels, const size_t* level_gates, const uint8_t* ops,
                                   const int32_t* in0, const int32_t* in1, const int32_t* out, size_t n_wires, size_t shard_min,
                                   tfhe_b200_group_circuit** circuit);
int tfhe_b200_group_circuit_run(tfhe_b200_group* g, tfhe_b200_group_circuit* circuit, const uint32_t* inputs_host, size_t n_inputs,
                                const int32_t* const_wires, const uint8_t* const_bits, size_t n_consts, const int32_t* out_wires,
                                size_t n_out, uint32_t* outputs_host);
int tfhe_b200_group_circuit_stats(const tfhe_b200_group_circuit* circuit, uint64_t* sharded_levels, uint64_t* replicated_levels,
                                  size_t* shard_min);
int tfhe_b200_group_circuit_destroy(tfhe_b200_group* g, tfhe_b200_group_circuit* circuit);
int tfhe_b200_host_alloc(void** out, size_t bytes);   /* pinned, portable across the devices of the box */
int tfhe_b200_host_free(void* p);

const char* tfhe_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif
