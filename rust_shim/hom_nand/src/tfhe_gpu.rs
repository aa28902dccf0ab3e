//! hom_nand/src/tfhe_gpu.rs -- how `TFHE` (hom_nand/src/tfhe.rs:9-113) keeps its public API on top of the C ABI.
//! `TFHE::new` still builds BK and KSK on the host exactly as today (tfhe.rs:21-25); it additionally flattens them and
//! hands them to the device context.  Single-gate methods call the batch entry with B = 1.
//! NOT COMPILED in this repository's image (no Rust toolchain); see INTEGRATION.md.
use crate::tlwe::TLWERep;
use utils::tfhe_b200 as ffi;

pub struct DeviceTfhe<const TLWE_N: usize> {
    ctx: *mut ffi::Ctx,
}

impl<const TLWE_N: usize> DeviceTfhe<TLWE_N> {
    /// bk_torus: [n][2l][2][N] torus words, rows in the order of TRGSWRep { cipher[2l], p_key[2l] } interleaved per row
    /// as (cipher[j], p_key[j]) (trgsw.rs:23-26); ksk: [N][t][3][n+1] = KeySwitchingKey::get(i, l, d) for d = 1..3.
    pub fn new(bk_torus: &[u32], ksk: &[u32]) -> Result<Self, String> {
        let mut ctx = std::ptr::null_mut();
        unsafe {
            check(ffi::tfhe_b200_ctx_create(std::ptr::null(), 0, &mut ctx), std::ptr::null())?;
            check(ffi::tfhe_b200_load_ksk(ctx, ksk.as_ptr()), ctx)?;
            check(ffi::tfhe_b200_load_bk(ctx, bk_torus.as_ptr()), ctx)?;
        }
        Ok(DeviceTfhe { ctx })
    }

    /// TLWERep is not repr(C): marshal explicitly to [b, a_0 .. a_{n-1}] (Torus32 is a newtype around u32 that the
    /// reference already pointer-casts to c_uint, utils/src/spqlios.rs:71-75).
    fn flatten(reps: &[TLWERep<TLWE_N>]) -> Vec<u32> {
        let mut v = Vec::with_capacity(reps.len() * (TLWE_N + 1));
        for r in reps {
            v.push(r.cipher().inner());
            v.extend(r.p_key().iter().map(|t| t.inner()));
        }
        v
    }

    pub fn gate_batch(&self, op: i32, in0: &[TLWERep<TLWE_N>], in1: &[TLWERep<TLWE_N>]) -> Result<Vec<TLWERep<TLWE_N>>, String> {
        let (a, b) = (Self::flatten(in0), Self::flatten(in1));
        let mut out = vec![0u32; a.len()];
        unsafe { check(ffi::tfhe_b200_gate_batch(self.ctx, op, a.as_ptr(), b.as_ptr(), out.as_mut_ptr(), in0.len()), self.ctx)?; }
        Ok(out.chunks(TLWE_N + 1).map(|c| TLWERep::from_words(c)).collect())
    }
    /// TFHE::hom_nand (tfhe.rs:41-47) keeps its signature: one gate = a batch of one.
    pub fn hom_nand(&self, input_0: TLWERep<TLWE_N>, input_1: TLWERep<TLWE_N>) -> TLWERep<TLWE_N> {
        self.gate_batch(ffi::NAND, &[input_0], &[input_1]).expect("tfhe_b200").pop().unwrap()
    }
}

impl<const TLWE_N: usize> Drop for DeviceTfhe<TLWE_N> {
    fn drop(&mut self) {
        unsafe { ffi::tfhe_b200_ctx_destroy(self.ctx); }
    }
}

unsafe fn check(rc: i32, ctx: *const ffi::Ctx) -> Result<(), String> {
    if rc == 0 { return Ok(()); }
    let msg = std::ffi::CStr::from_ptr(ffi::tfhe_b200_last_error(ctx)).to_string_lossy().into_owned();
    Err(format!("tfhe_b200 error {rc}: {msg}"))
}
