//! utils/src/tfhe_b200.rs -- the `extern "C"` block that replaces utils/src/spqlios.rs:18-32.
//! One call per BATCH of gates instead of 8 FFT entry points crossed 5080 times per gate.
//! Mirrors include/tfhe_b200.h one to one.  NOT COMPILED in this repository's image (no Rust toolchain).
use std::os::raw::{c_char, c_int, c_void};

pub enum Ctx {}
pub enum Circuit {}
pub enum Group {}
pub enum GroupCircuit {}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct Params {
    pub n: i32,
    pub big_n: i32,
    pub l: i32,
    pub bgbit: i32,
    pub ks_t: i32,
    pub ks_basebit: i32,
    pub mu: u32,
    pub decomp_mask: u32,
}

pub const NAND: c_int = 0;
pub const AND: c_int = 1;
pub const OR: c_int = 2;
pub const XOR: c_int = 3;
pub const NOT: c_int = 4;
pub const COPY: c_int = 5;
pub const ANDNY: c_int = 6;
// file kinds of tfhe_b200_file_{write,info,read}
pub const FILE_SECRET: c_int = 1;
pub const FILE_BK: c_int = 2;
pub const FILE_KSK: c_int = 3;
pub const FILE_TLWE0: c_int = 4;
pub const FILE_TLWE1: c_int = 5;
pub const FILE_TRLWE: c_int = 6;
pub const FILE_TRGSW: c_int = 7;

extern "C" {
    pub fn tfhe_b200_default_params(p: *mut Params) -> c_int;
    pub fn tfhe_b200_ctx_create(p: *const Params, device: c_int, out: *mut *mut Ctx) -> c_int;
    pub fn tfhe_b200_ctx_destroy(ctx: *mut Ctx) -> c_int;
    pub fn tfhe_b200_last_error(ctx: *const Ctx) -> *const c_char;
    pub fn tfhe_b200_load_bk(ctx: *mut Ctx, bk: *const u32) -> c_int;
    pub fn tfhe_b200_load_ksk(ctx: *mut Ctx, ksk: *const u32) -> c_int;
    pub fn tfhe_b200_gate_batch(ctx: *mut Ctx, op: c_int, in0: *const u32, in1: *const u32, out: *mut u32, b: usize) -> c_int;
    pub fn tfhe_b200_gate_batch_device(ctx: *mut Ctx, op: c_int, in0: *const u32, in1: *const u32, out: *mut u32, b: usize, stream: *mut c_void) -> c_int;
    pub fn tfhe_b200_mux_batch(ctx: *mut Ctx, c: *const u32, in0: *const u32, in1: *const u32, out: *mut u32, b: usize) -> c_int;
    pub fn tfhe_b200_blind_rotate_batch(ctx: *mut Ctx, input: *const u32, nsteps: c_int, out_trlwe: *mut u32, b: usize) -> c_int;
    pub fn tfhe_b200_keyswitch_batch(ctx: *mut Ctx, lwe1: *const u32, out: *mut u32, b: usize) -> c_int;
    pub fn tfhe_b200_external_product_batch(ctx: *mut Ctx, trgsw: *const u32, ntrgsw: usize, trlwe: *const u32, out: *mut u32, b: usize) -> c_int;
    pub fn tfhe_b200_negacyclic_mul_batch(ctx: *mut Ctx, a: *const u32, d: *const i32, out: *mut u32, b: usize) -> c_int;
    // asynchronous host-pointer form + workspace reservation
    pub fn tfhe_b200_gate_batch_async(ctx: *mut Ctx, op: c_int, in0: *const u32, in1: *const u32, out: *mut u32, b: usize) -> c_int;
    pub fn tfhe_b200_sync(ctx: *mut Ctx) -> c_int;
    pub fn tfhe_b200_gate_batch_mixed(ctx: *mut Ctx, ops: *const u8, in0: *const u32, in1: *const u32, out: *mut u32, b: usize) -> c_int;
    pub fn tfhe_b200_set_key_slices(ctx: *mut Ctx, slices: c_int) -> c_int;
    pub fn tfhe_b200_set_batch_overlap(ctx: *mut Ctx, mode: c_int) -> c_int;
    pub fn tfhe_b200_reserve(ctx: *mut Ctx, max_batch: usize) -> c_int;
    pub fn tfhe_b200_bootstrap_batch(ctx: *mut Ctx, input: *const u32, out: *mut u32, b: usize) -> c_int;
    pub fn tfhe_b200_bootstrap_lv1_batch(ctx: *mut Ctx, input: *const u32, out_lwe1: *mut u32, b: usize) -> c_int;
    pub fn tfhe_b200_cmux_batch(ctx: *mut Ctx, trgsw: *const u32, ntrgsw: usize, rep1: *const u32, rep0: *const u32, out: *mut u32, b: usize) -> c_int;
    pub fn tfhe_b200_sample_extract_batch(ctx: *mut Ctx, trlwe: *const u32, index: c_int, out_lwe1: *mut u32, b: usize) -> c_int;
    // device-side key generation (replaces the host loops of BootstrappingKey::new / KeySwitchingKey::new), export, encryption
    pub fn tfhe_b200_keygen_device(ctx: *mut Ctx, seed: u64, s0: *const u8, s1: *const u8) -> c_int;
    pub fn tfhe_b200_export_bk(ctx: *mut Ctx, bk: *mut u32) -> c_int;
    pub fn tfhe_b200_export_ksk(ctx: *mut Ctx, ksk: *mut u32) -> c_int;
    pub fn tfhe_b200_encrypt_bits_device(ctx: *mut Ctx, seed: u64, ct_index0: u64, s0: *const u8, bits_dev: *const u8, b: usize, out_dev: *mut u32, stream: *mut c_void) -> c_int;
    pub fn tfhe_b200_decrypt_bits_device(ctx: *mut Ctx, s0: *const u8, ct_dev: *const u32, b: usize, bits_dev: *mut u8, phase_dev: *mut u32, stream: *mut c_void) -> c_int;
    // device-resident circuits (the level-synchronous evaluator behind nander's LogicExpr, nander/src/lib.rs:72-89)
    pub fn tfhe_b200_circuit_create(ctx: *mut Ctx, n_levels: usize, level_gates: *const usize, ops: *const u8, in0: *const i32, in1: *const i32,
                                    out: *const i32, n_wires: usize, circuit: *mut *mut Circuit) -> c_int;
    pub fn tfhe_b200_circuit_run_device(ctx: *mut Ctx, circuit: *const Circuit, wires_dev: *mut u32, stream: *mut c_void) -> c_int;
    pub fn tfhe_b200_circuit_destroy(ctx: *mut Ctx, circuit: *mut Circuit) -> c_int;
    // several GPUs from one process: keys replicated by an NCCL broadcast, batches sharded, circuits with a per-level exchange
    pub fn tfhe_b200_group_create(p: *const Params, devices: *const c_int, ndev: c_int, out: *mut *mut Group) -> c_int;
    pub fn tfhe_b200_group_destroy(g: *mut Group) -> c_int;
    pub fn tfhe_b200_group_last_error(g: *const Group) -> *const c_char;
    pub fn tfhe_b200_group_size(g: *const Group) -> c_int;
    pub fn tfhe_b200_group_load_bk(g: *mut Group, bk: *const u32) -> c_int;
    pub fn tfhe_b200_group_load_ksk(g: *mut Group, ksk: *const u32) -> c_int;
    pub fn tfhe_b200_group_keygen_csprng(g: *mut Group, key: *const u8, s0: *const u8, s1: *const u8) -> c_int;
    pub fn tfhe_b200_group_reserve(g: *mut Group, max_batch: usize) -> c_int;
    pub fn tfhe_b200_group_gate_batch(g: *mut Group, op: c_int, in0: *const u32, in1: *const u32, out: *mut u32, b: usize) -> c_int;
    pub fn tfhe_b200_group_gate_batch_async(g: *mut Group, op: c_int, in0: *const u32, in1: *const u32, out: *mut u32, b: usize) -> c_int;
    pub fn tfhe_b200_group_sync(g: *mut Group) -> c_int;
    pub fn tfhe_b200_group_circuit_create(g: *mut Group, n_levels: usize, level_gates: *const usize, ops: *const u8, in0: *const i32,
                                          in1: *const i32, out: *const i32, n_wires: usize, shard_min: usize,
                                          circuit: *mut *mut GroupCircuit) -> c_int;
    pub fn tfhe_b200_group_circuit_run(g: *mut Group, circuit: *mut GroupCircuit, inputs: *const u32, n_inputs: usize, const_wires: *const i32,
                                       const_bits: *const u8, n_consts: usize, out_wires: *const i32, n_out: usize, outputs: *mut u32) -> c_int;
    pub fn tfhe_b200_group_circuit_destroy(g: *mut Group, circuit: *mut GroupCircuit) -> c_int;
    pub fn tfhe_b200_host_alloc(out: *mut *mut c_void, bytes: usize) -> c_int;
    pub fn tfhe_b200_host_free(p: *mut c_void) -> c_int;
    // flat file format (the reference has no serialisation)
    pub fn tfhe_b200_file_write(path: *const c_char, kind: c_int, payload: *const c_void, count: u64) -> c_int;
    pub fn tfhe_b200_file_info(path: *const c_char, kind: *mut c_int, count: *mut u64, payload_bytes: *mut u64) -> c_int;
    pub fn tfhe_b200_file_read(path: *const c_char, kind: c_int, payload: *mut c_void, payload_bytes: u64) -> c_int;
    pub fn tfhe_b200_file_last_error() -> *const c_char;
}
