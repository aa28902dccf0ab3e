"""ctypes loader for the CPU oracle (oracle/liboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
The product package (rustfhe_b200) never does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
n, N, L, KS_T = 635, 1024, 3, 8
MU = 0x20000000
MASK_FAITHFUL = 0x02084000
MASK_TESTED = 0x02082000
NAND, AND, OR, XOR, NOT, COPY, ANDNY = range(7)
BK_WORDS = n * 2 * L * 2 * N
KSK_WORDS = N * KS_T * 3 * (n + 1)

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(ref=True):
    """Compile liboracle.so (and _ref/libspqlios_ref.so when /root/reference is present). Building the checker is
    not using it."""
    subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    if ref and os.path.isdir("/root/reference/utils/src/spqlios"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(HERE, "liboracle.so")
    if not os.path.exists(path):
        build(ref=False)
    l = C.CDLL(path)
    sz, u64, u32, i32, dbl, vp = C.c_size_t, C.c_uint64, C.c_uint32, C.c_int, C.c_double, C.c_void_p
    sig = {
        "orc_rnd64": (u64, [u64, u64, u64]),
        "orc_gauss_torus": (C.c_int32, [u64, u64, u64, dbl]),
        "orc_keygen_secret": (None, [u64, u8p, u8p]),
        "orc_keygen_bk": (None, [u64, u8p, u8p, u32p]),
        "orc_keygen_ksk": (None, [u64, u8p, u8p, u32p]),
        "orc_tlwe_encrypt_bits": (None, [u64, u64, u8p, u8p, sz, u32p]),
        "orc_tlwe_phase": (None, [u8p, u32p, sz, u32p]),
        "orc_tlwe_decrypt_bits": (None, [u8p, u32p, sz, u8p]),
        "orc_tlwe1_phase": (None, [u8p, u32p, sz, u32p]),
        "orc_trlwe_phase": (None, [u8p, u32p, u32p]),
        "orc_rotate": (None, [u32p, i32, i32, u32p]),
        "orc_negacyclic_mul_schoolbook": (None, [u32p, i32p, i32, u32p]),
        "orc_make_decomp_mask": (u32, [u32, u32]),
        "orc_tested_decomp_mask": (u32, [u32, u32]),
        "orc_decompose_scalar": (None, [u32, u32, u32, u32, i32p]),
        "orc_decompose": (None, [u32p, u32, i32p]),
        "orc_decompose_u32_scalar": (None, [u32, u32, u32, u32p]),
        "orc_ref_roundtrip_n": (None, [i32, u32p, u32p]),
        "orc_ref_poly_mul_n": (None, [i32, u32p, u32p, u32p]),
        "orc_torus_from_f32": (u32, [C.c_float]),
        "orc_torus_to_f32": (C.c_float, [u32]),
        "orc_gate_linear": (None, [i32, u32p, vp, sz, u32p]),
        "orc_sample_extract0": (None, [u32p, u32p]),
        "orc_sample_extract": (None, [u32p, i32, u32p]),
        "orc_key_switch": (None, [u32p, u32p, u32p]),
        "orc_ks_digits": (None, [u32p, u16p]),
        "orc_negacyclic_mul_ntt": (None, [u32p, i32p, u32p]),
        "orc_external_product_exact": (None, [u32p, u32p, u32, u32p]),
        "orc_exact_bk_prepare": (vp, [u32p]),
        "orc_exact_bk_free": (None, [vp]),
        "orc_blind_rotate_exact": (None, [vp, u32p, u32, i32, u32p]),
        "orc_bootstrap_exact": (None, [vp, u32p, u32p, sz, u32, u32p, vp]),
        "orc_ref_init": (i32, [C.c_char_p]),
        "orc_ref_available": (i32, []),
        "orc_ref_poly_mul": (None, [u32p, u32p, u32p]),
        "orc_ref_ifft_fft_roundtrip": (None, [u32p, u32p]),
        "orc_ref_bk_fourier": (vp, [u32p]),
        "orc_ref_free": (None, [vp]),
        "orc_ref_external_product_torus": (None, [u32p, u32p, u32, u32p]),
        "orc_ref_blind_rotate": (None, [vp, u32p, u32, i32, u32p]),
        "orc_ref_bootstrap": (None, [vp, u32p, u32p, sz, u32, i32, u32p]),
        "orc_ref_bench_gates": (dbl, [vp, u32p, u32p, vp, i32, sz, u32, i32, u32p]),
    }
    for name, (res, args) in sig.items():
        f = getattr(l, name)
        f.restype, f.argtypes = res, args
    _lib = l
    return l


def ref_path():
    return os.path.join(HERE, "_ref", "libspqlios_ref.so")


def ref_init():
    """Load the reference's own FFT library (oracle/_ref). Returns True when available."""
    if not os.path.exists(ref_path()):
        return False
    return lib().orc_ref_init(ref_path().encode()) == 0


class Keys:
    """Seeded key material as flat little-endian u32 buffers (layouts: include/tfhe_b200.h)."""

    def __init__(self, seed=0x5EED0001):
        l = lib()
        self.seed = seed
        self.s0 = np.zeros(n, np.uint8)
        self.s1 = np.zeros(N, np.uint8)
        l.orc_keygen_secret(seed, self.s0, self.s1)
        self.bk = np.zeros(BK_WORDS, np.uint32)
        self.ksk = np.zeros(KSK_WORDS, np.uint32)
        l.orc_keygen_bk(seed, self.s0, self.s1, self.bk)
        l.orc_keygen_ksk(seed, self.s0, self.s1, self.ksk)
        self._exact = None
        self._fourier = None

    def exact_handle(self):
        if self._exact is None:
            self._exact = lib().orc_exact_bk_prepare(self.bk)
        return self._exact

    def fourier_handle(self):
        if self._fourier is None:
            assert ref_init(), "oracle/_ref/libspqlios_ref.so missing"
            self._fourier = lib().orc_ref_bk_fourier(self.bk)
        return self._fourier

    def encrypt(self, bits, ct_index0=0, seed=None):
        bits = np.ascontiguousarray(bits, np.uint8)
        out = np.zeros((len(bits), n + 1), np.uint32)
        lib().orc_tlwe_encrypt_bits(self.seed + 1 if seed is None else seed, ct_index0, self.s0, bits, len(bits), out)
        return out

    def phase(self, ct):
        ct = np.ascontiguousarray(ct, np.uint32).reshape(-1, n + 1)
        ph = np.zeros(len(ct), np.uint32)
        lib().orc_tlwe_phase(self.s0, ct, len(ct), ph)
        return ph

    def decrypt(self, ct):
        ct = np.ascontiguousarray(ct, np.uint32).reshape(-1, n + 1)
        bits = np.zeros(len(ct), np.uint8)
        lib().orc_tlwe_decrypt_bits(self.s0, ct, len(ct), bits)
        return bits


def gate_linear(op, in0, in1=None):
    in0 = np.ascontiguousarray(in0, np.uint32).reshape(-1, n + 1)
    out = np.zeros_like(in0)
    p1 = None
    if in1 is not None:
        in1 = np.ascontiguousarray(in1, np.uint32).reshape(-1, n + 1)
        p1 = in1.ctypes.data_as(C.c_void_p)
    lib().orc_gate_linear(op, in0, p1, len(in0), out)
    return out


def bootstrap_exact(keys, lin, mask=MASK_FAITHFUL, want_lwe1=False):
    lin = np.ascontiguousarray(lin, np.uint32).reshape(-1, n + 1)
    out = np.zeros_like(lin)
    lwe1 = np.zeros((len(lin), N + 1), np.uint32) if want_lwe1 else None
    lib().orc_bootstrap_exact(keys.exact_handle(), keys.ksk, lin, len(lin), mask, out,
                              lwe1.ctypes.data_as(C.c_void_p) if want_lwe1 else None)
    return (out, lwe1) if want_lwe1 else out


def bootstrap_ref(keys, lin, mask=MASK_FAITHFUL, nthreads=0):
    lin = np.ascontiguousarray(lin, np.uint32).reshape(-1, n + 1)
    out = np.zeros_like(lin)
    lib().orc_ref_bootstrap(keys.fourier_handle(), keys.ksk, lin, len(lin), mask, nthreads or (os.cpu_count() or 1), out)
    return out


def gate_exact(keys, op, in0, in1=None, mask=MASK_FAITHFUL):
    return bootstrap_exact(keys, gate_linear(op, in0, in1), mask)


def gate_ref(keys, op, in0, in1=None, mask=MASK_FAITHFUL, nthreads=0):
    return bootstrap_ref(keys, gate_linear(op, in0, in1), mask, nthreads)


def bench_ref_gates(keys, op, in0, in1, nthreads=1, mask=MASK_FAITHFUL):
    """Wall seconds for len(in0) bootstrapped gates on `nthreads` host threads (reference FFT + restated glue)."""
    in0 = np.ascontiguousarray(in0, np.uint32).reshape(-1, n + 1)
    out = np.zeros_like(in0)
    p1 = None
    if in1 is not None:
        in1 = np.ascontiguousarray(in1, np.uint32).reshape(-1, n + 1)
        p1 = in1.ctypes.data_as(C.c_void_p)
    secs = lib().orc_ref_bench_gates(keys.fourier_handle(), keys.ksk, in0, p1, op, len(in0), mask, nthreads, out)
    return secs, out
