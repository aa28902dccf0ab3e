//! hom_nand/src/tfhe.rs -- DROP-IN REPLACEMENT for the reference file of the same name (hom_nand/src/tfhe.rs:1-135).
//!
//! Same public items, same signatures: `TFHE::new`, `hom_mux / hom_nand / hom_and / hom_or / hom_xor / hom_not`,
//! `TFHEHelper::{NBIT, COEF}`, `BootstrappingKey::{new, iter}` -- so `nander`'s `impl Logip for TFHE` (nander/src/lib.rs:40-62)
//! and `examples/homnand-bench.rs` compile unchanged.  What changes is what happens inside: `bootstrap` (blind rotation
//! of 635 CMUXes, sample extract, key switch; tfhe.rs:73-113) is ONE call into the CUDA engine (include/tfhe_b200.h) instead
//! of 5080 FFT calls through utils/src/spqlios.rs, and every gate also exists as a `*_batch` method (B independent gates, one
//! launch pair).  Keys are still generated on the host by the reference's own `KeySwitchingKey::new` and
//! `Cryptor::encrypto(TRGSW, ..)`; they are flattened into the ABI layouts here and uploaded once.
//!
//! Every reference item used below exists today:
//!   digest::{Cryptor, Encrypted}                       hom_nand/src/digest.rs:14-42
//!   tlwe::{KeySwitchingKey::{new, get}, TLWEHelper::{IKS_L, N}, TLWERep::{new, cipher, p_key}}   tlwe.rs:19-41, 175-180, 243-283
//!   trgsw::{TRGSW, TRGSWHelper::L, TRGSWRep::{cipher, p_key}}                                    trgsw.rs:11-62, 110-116
//!   utils::math::{Binary, Polynomial::{new, coefs}, Torus32::{from_bits, inner}}                 math.rs:42-52, 362-366, 490-500
//! NOT COMPILED in this repository's image (no cargo / rustc; the crate needs nightly features).
use crate::digest::{Cryptor, Encrypted};
use crate::tlwe::{KeySwitchingKey, TLWEHelper, TLWERep};
use crate::trgsw::{TRGSWHelper, TRGSWRep, TRGSW};
use utils::math::{Binary, Polynomial, Torus32};
use utils::pol;
use utils::tfhe_b200 as ffi;

pub struct TFHE<const TLWE_N: usize, const TRLWE_N: usize> {
    ctx: *mut ffi::Ctx,
    /// the torus-domain keys stay available to callers that used to read them (`BootstrappingKey::iter`)
    bk: BootstrappingKey<TLWE_N, TRLWE_N>,
}

pub struct TFHEHelper;
impl TFHEHelper {
    pub const NBIT: u32 = 10; // = log_2(TRLWEHelper::N)
    pub const COEF: f32 = 1. / 8.;
}

/// gate opcodes of the C ABI (include/tfhe_b200.h); the linear pre-combination of tfhe.rs:27-71 runs in the kernel prologue
#[derive(Clone, Copy)]
enum Op {
    Nand = ffi::NAND as isize,
    And = ffi::AND as isize,
    Or = ffi::OR as isize,
    Xor = ffi::XOR as isize,
    Not = ffi::NOT as isize,
}

impl<const TLWE_N: usize, const TRLWE_N: usize> TFHE<TLWE_N, TRLWE_N> {
    /// tfhe.rs:21-25.  The engine is built for the crate's default parameter set only (n = 635, N = 1024, l = 3, Bg = 64,
    /// t = 8, basebit = 2); other const parameters panic here instead of computing something else.
    pub fn new(s_key_tlwelv0: [Binary; TLWE_N], s_key_tlwelv1: [Binary; TRLWE_N]) -> Self {
        assert!(TLWE_N == TLWEHelper::N && TRLWE_N == 1 << TFHEHelper::NBIT, "tfhe_b200 supports TLWE_N = 635, TRLWE_N = 1024");
        let ksk = KeySwitchingKey::new(s_key_tlwelv1, &s_key_tlwelv0);
        let bk = BootstrappingKey::new(s_key_tlwelv0, &pol!(s_key_tlwelv1));
        let ksk_words = flatten_ksk::<TRLWE_N, TLWE_N>(&ksk);
        let mut ctx = std::ptr::null_mut();
        unsafe {
            check(ffi::tfhe_b200_ctx_create(std::ptr::null(), 0, &mut ctx), std::ptr::null());
            check(ffi::tfhe_b200_load_ksk(ctx, ksk_words.as_ptr()), ctx);
            check(ffi::tfhe_b200_load_bk(ctx, bk.words().as_ptr()), ctx);
        }
        TFHE { ctx, bk }
    }
    /// (input_1&control)|(input_0&!control)   -- tfhe.rs:27-40; both first-stage ANDs share one launch in the engine
    pub fn hom_mux(&self, control: TLWERep<TLWE_N>, input_0: TLWERep<TLWE_N>, input_1: TLWERep<TLWE_N>) -> TLWERep<TLWE_N> {
        self.hom_mux_batch(&[control], &[input_0], &[input_1]).pop().unwrap()
    }
    pub fn hom_nand(&self, input_0: TLWERep<TLWE_N>, input_1: TLWERep<TLWE_N>) -> TLWERep<TLWE_N> {
        self.gate_batch(Op::Nand, &[input_0], Some(&[input_1])).pop().unwrap() // tfhe.rs:41-47
    }
    pub fn hom_and(&self, input_0: TLWERep<TLWE_N>, input_1: TLWERep<TLWE_N>) -> TLWERep<TLWE_N> {
        self.gate_batch(Op::And, &[input_0], Some(&[input_1])).pop().unwrap() // tfhe.rs:48-54
    }
    pub fn hom_or(&self, input_0: TLWERep<TLWE_N>, input_1: TLWERep<TLWE_N>) -> TLWERep<TLWE_N> {
        self.gate_batch(Op::Or, &[input_0], Some(&[input_1])).pop().unwrap() // tfhe.rs:55-61
    }
    pub fn hom_xor(&self, input_0: TLWERep<TLWE_N>, input_1: TLWERep<TLWE_N>) -> TLWERep<TLWE_N> {
        self.gate_batch(Op::Xor, &[input_0], Some(&[input_1])).pop().unwrap() // tfhe.rs:62-68
    }
    pub fn hom_not(&self, input: TLWERep<TLWE_N>) -> TLWERep<TLWE_N> {
        self.gate_batch(Op::Not, &[input], None).pop().unwrap() // tfhe.rs:69-71 (still bootstrapped)
    }

    // ---- the batch forms: B independent gates, one blind-rotation launch + one key-switch launch ----
    pub fn hom_nand_batch(&self, input_0: &[TLWERep<TLWE_N>], input_1: &[TLWERep<TLWE_N>]) -> Vec<TLWERep<TLWE_N>> {
        self.gate_batch(Op::Nand, input_0, Some(input_1))
    }
    pub fn hom_and_batch(&self, input_0: &[TLWERep<TLWE_N>], input_1: &[TLWERep<TLWE_N>]) -> Vec<TLWERep<TLWE_N>> {
        self.gate_batch(Op::And, input_0, Some(input_1))
    }
    pub fn hom_or_batch(&self, input_0: &[TLWERep<TLWE_N>], input_1: &[TLWERep<TLWE_N>]) -> Vec<TLWERep<TLWE_N>> {
        self.gate_batch(Op::Or, input_0, Some(input_1))
    }
    pub fn hom_xor_batch(&self, input_0: &[TLWERep<TLWE_N>], input_1: &[TLWERep<TLWE_N>]) -> Vec<TLWERep<TLWE_N>> {
        self.gate_batch(Op::Xor, input_0, Some(input_1))
    }
    pub fn hom_not_batch(&self, input: &[TLWERep<TLWE_N>]) -> Vec<TLWERep<TLWE_N>> {
        self.gate_batch(Op::Not, input, None)
    }
    pub fn hom_mux_batch(&self, control: &[TLWERep<TLWE_N>], input_0: &[TLWERep<TLWE_N>], input_1: &[TLWERep<TLWE_N>]) -> Vec<TLWERep<TLWE_N>> {
        assert!(control.len() == input_0.len() && control.len() == input_1.len(), "hom_mux_batch: operand batches differ in length");
        let (c, a, b) = (flatten_tlwe(control), flatten_tlwe(input_0), flatten_tlwe(input_1));
        let mut out = vec![0u32; c.len()];
        unsafe { check(ffi::tfhe_b200_mux_batch(self.ctx, c.as_ptr(), a.as_ptr(), b.as_ptr(), out.as_mut_ptr(), control.len()), self.ctx) };
        unflatten_tlwe(&out)
    }
    pub fn bootstrapping_key(&self) -> &BootstrappingKey<TLWE_N, TRLWE_N> {
        &self.bk
    }

    fn gate_batch(&self, op: Op, input_0: &[TLWERep<TLWE_N>], input_1: Option<&[TLWERep<TLWE_N>]>) -> Vec<TLWERep<TLWE_N>> {
        if let Some(b) = input_1 {
            assert!(b.len() == input_0.len(), "hom_*_batch: operand batches differ in length");
        }
        let a = flatten_tlwe(input_0);
        let b = input_1.map(flatten_tlwe);
        let mut out = vec![0u32; a.len()];
        let pb = b.as_ref().map_or(std::ptr::null(), |v| v.as_ptr());
        unsafe { check(ffi::tfhe_b200_gate_batch(self.ctx, op as i32, a.as_ptr(), pb, out.as_mut_ptr(), input_0.len()), self.ctx) };
        unflatten_tlwe(&out)
    }
}

impl<const TLWE_N: usize, const TRLWE_N: usize> Drop for TFHE<TLWE_N, TRLWE_N> {
    fn drop(&mut self) {
        unsafe { ffi::tfhe_b200_ctx_destroy(self.ctx) };
    }
}

/// TLWERep is not repr(C): marshal explicitly to the ABI layout [b, a_0 .. a_{n-1}] (tlwe.rs:19-23: cipher, p_key)
fn flatten_tlwe<const N: usize>(reps: &[TLWERep<N>]) -> Vec<u32> {
    let mut v = Vec::with_capacity(reps.len() * (N + 1));
    for r in reps {
        v.push(r.cipher().inner());
        v.extend(r.p_key().iter().map(|t| t.inner()));
    }
    v
}
fn unflatten_tlwe<const N: usize>(words: &[u32]) -> Vec<TLWERep<N>> {
    words
        .chunks_exact(N + 1)
        .map(|w| {
            let mut p_key = [Torus32::from_bits(0); N];
            for (d, s) in p_key.iter_mut().zip(&w[1..]) {
                *d = Torus32::from_bits(*s);
            }
            TLWERep::new(Torus32::from_bits(w[0]), p_key)
        })
        .collect()
}
/// KeySwitchingKey (tlwe.rs:243-283: Vec<[[TLWERep<M>; IKS_T]; IKS_L]>, get(i, l, t) = KS[i][l][t-1]) -> ABI layout
/// [N][t = IKS_L][3][M + 1]: only t = 1..3 is reachable (a 2-bit digit, tlwe.rs:58-69); the stored t = 4 entry is dropped.
fn flatten_ksk<const N: usize, const M: usize>(ksk: &KeySwitchingKey<N, M>) -> Vec<u32> {
    let mut v = Vec::with_capacity(N * TLWEHelper::IKS_L * 3 * (M + 1));
    for i in 0..N {
        for l in 0..TLWEHelper::IKS_L {
            for t in 1..=3usize {
                let r = ksk.get(i, l, t);
                v.push(r.cipher().inner());
                v.extend(r.p_key().iter().map(|x| x.inner()));
            }
        }
    }
    v
}

/// tfhe.rs:116-135 keeps a Vec<TRGSWRepF> (Fourier domain).  The engine transforms the key itself (into its NTT domain, on the
/// device), so this version keeps the TORUS-domain samples and their flat image: [n][2l][2][N] u32, rows in the order of
/// TRGSWRep { cipher[2l], p_key[2l] } (trgsw.rs:23-26), per row the `cipher` polynomial first, then `p_key`.
pub struct BootstrappingKey<const PRE_N: usize, const N: usize> {
    reps: Vec<TRGSWRep<N>>,
    words: Vec<u32>,
}
impl<const PRE_N: usize, const N: usize> BootstrappingKey<PRE_N, N> {
    pub fn new(s_key_tlwe: [Binary; PRE_N], s_key: &Polynomial<Binary, N>) -> Self {
        let mut reps = Vec::<TRGSWRep<N>>::with_capacity(PRE_N);
        let mut words = Vec::<u32>::with_capacity(PRE_N * 2 * TRGSWHelper::L * 2 * N);
        for s_i in s_key_tlwe {
            let trgsw_: TRGSWRep<N> = Cryptor::encrypto(TRGSW, s_key, s_i); // tfhe.rs:122, trgsw.rs:250-256
            for j in 0..2 * TRGSWHelper::L {
                words.extend(trgsw_.cipher()[j].coefs().iter().map(|t| t.inner()));
                words.extend(trgsw_.p_key()[j].coefs().iter().map(|t| t.inner()));
            }
            reps.push(trgsw_);
        }
        BootstrappingKey { reps, words }
    }
    #[inline]
    pub fn iter(&self) -> std::slice::Iter<'_, TRGSWRep<N>> {
        self.reps.iter()
    }
    /// the flat image handed to `tfhe_b200_load_bk`
    pub fn words(&self) -> &[u32] {
        &self.words
    }
}

unsafe fn check(rc: i32, ctx: *const ffi::Ctx) {
    if rc != 0 {
        let msg = std::ffi::CStr::from_ptr(ffi::tfhe_b200_last_error(ctx)).to_string_lossy().into_owned();
        // the reference aborts on precondition failures (spqlios-fft-impl.cpp:92-97); there is no CPU fallback to fall to
        panic!("tfhe_b200 error {}: {}", rc, msg);
    }
}
