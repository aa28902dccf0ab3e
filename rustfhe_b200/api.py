"""Host-side mirror of the reference's hom_nand API for the bootstrapped-gate path, over the C ABI.

Reference interface mirrored here (file:line relative to /root/reference):
  TFHE::{new, hom_nand, hom_and, hom_or, hom_xor, hom_not, hom_mux}      hom_nand/src/tfhe.rs:21-71
  TFHEHelper / TLWEHelper / TRLWEHelper / TRGSWHelper constants           tfhe.rs:14-18, tlwe.rs:173-195, trlwe.rs:74-78, trgsw.rs:111-116
  TLWERep::{new, trivial, +, -, neg, *, identity_key_switch}             hom_nand/src/tlwe.rs:19-159
  KeySwitchingKey::new, BootstrappingKey::new                            tlwe.rs:247-277, tfhe.rs:119-126
  Cryptor::{encrypto, decrypto} with the TLWE strategy on Binary          digest.rs:14-33, tlwe.rs:197-241
Every gate method takes BATCHES (arrays [B][n+1]); a single ciphertext is a batch of one.
"""
import ctypes as C

import numpy as np

from . import _capi as K
from ._capi import TfheError, lib, ptr


class TFHEHelper:            # tfhe.rs:14-18
    NBIT = 10
    COEF = 1.0 / 8.0


class TLWEHelper:            # tlwe.rs:173-195
    N = 635
    ALPHA = 2.0 ** -15
    IKS_L = 8
    IKS_BASEBIT = 2
    IKS_T = 4

    @staticmethod
    def binary2torus(bit):
        return np.uint32(0x20000000) if bit else np.uint32(0xE0000000)

    @staticmethod
    def torus2binary(torus):
        return (np.asarray(torus, np.uint32).astype(np.float32) * np.float32(2.0 ** -32) < np.float32(0.5)).astype(np.uint8)


class TRLWEHelper:           # trlwe.rs:74-78
    N = 1024
    ALPHA = 2.0 ** -25


class TRGSWHelper:           # trgsw.rs:111-116
    BGBIT = 6
    BG = 64
    L = 3


class TLWE:
    """Strategy marker, as in `Cryptor::encrypto(TLWE, &s_key, item)` (tlwe.rs:10)."""


def _u32_batch(a, width):
    a = np.ascontiguousarray(a, dtype=np.uint32)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.ndim != 2 or a.shape[1] != width:
        raise ValueError(f"expected uint32 array [B][{width}], got {a.shape}")
    return a


class TLWERep:
    """Linear operations on level-0 ciphertext batches (tlwe.rs:76-159). Plain wrapping u32 arithmetic, host side."""

    @staticmethod
    def new(cipher, p_key):
        return np.concatenate([np.asarray(cipher, np.uint32).reshape(-1, 1), _u32_batch(p_key, K.n)], axis=1)

    @staticmethod
    def trivial(text, batch=1):   # tlwe.rs:76-79
        out = np.zeros((batch, K.n + 1), np.uint32)
        out[:, 0] = np.uint32(text)
        return out

    @staticmethod
    def logic_true(batch=1):      # AsLogic, tlwe.rs:80-87
        return TLWERep.trivial(0x20000000, batch)

    @staticmethod
    def logic_false(batch=1):
        return TLWERep.trivial(0xE0000000, batch)

    @staticmethod
    def add(x, y):
        return (_u32_batch(x, K.n + 1) + _u32_batch(y, K.n + 1)).astype(np.uint32)

    @staticmethod
    def sub(x, y):
        return (_u32_batch(x, K.n + 1) - _u32_batch(y, K.n + 1)).astype(np.uint32)

    @staticmethod
    def neg(x):
        return (np.uint32(0) - _u32_batch(x, K.n + 1)).astype(np.uint32)

    @staticmethod
    def mul(x, k):
        return (_u32_batch(x, K.n + 1) * np.uint32(int(k) & 0xFFFFFFFF)).astype(np.uint32)


class SecretKeys:
    """The two binary secret keys the reference samples with BinaryDistribution::uniform (homnand-bench.rs:10-12)."""

    def __init__(self, s_key_tlwelv0, s_key_tlwelv1):
        self.s_key_tlwelv0 = np.ascontiguousarray(s_key_tlwelv0, np.uint8)
        self.s_key_tlwelv1 = np.ascontiguousarray(s_key_tlwelv1, np.uint8)
        assert self.s_key_tlwelv0.shape == (K.n,) and self.s_key_tlwelv1.shape == (K.N,)

    @staticmethod
    def generate(seed=None):
        """seed=None (default): ChaCha20 keyed from getrandom(2), like the reference's thread_rng.  An integer seed selects the
        DETERMINISTIC TEST generator (reproducible, NOT secure -- parity tests only; see tfhe_rng.cuh)."""
        s0, s1 = np.zeros(K.n, np.uint8), np.zeros(K.N, np.uint8)
        if seed is None:
            _check(None, lib().tfhe_b200_keygen_secret_csprng(None, ptr(s0), ptr(s1)))
        else:
            _check(None, lib().tfhe_b200_keygen_secret(seed, ptr(s0), ptr(s1)))
        return SecretKeys(s0, s1)


def _check(ctx, rc):
    if rc != K.OK:
        msg = lib().tfhe_b200_last_error(ctx)
        raise TfheError(rc, msg.decode() if msg else "")


class KeySwitchingKey:
    """KeySwitchingKey::new(pre_s_key = lv1 key, next_s_key = lv0 key) (tlwe.rs:247-277); flat [N][t][3][n+1]."""

    def __init__(self, words):
        self.words = np.ascontiguousarray(words, np.uint32).reshape(-1)
        assert self.words.size == K.KSK_WORDS

    @staticmethod
    def new(s_key_tlwelv1, s_key_tlwelv0, seed=None):
        """seed=None: fresh CSPRNG key per call; integer seed: deterministic test generator (INSECURE)."""
        w = np.zeros(K.KSK_WORDS, np.uint32)
        s0p, s1p = ptr(np.ascontiguousarray(s_key_tlwelv0, np.uint8)), ptr(np.ascontiguousarray(s_key_tlwelv1, np.uint8))
        if seed is None:
            _check(None, lib().tfhe_b200_keygen_ksk_csprng(None, s0p, s1p, ptr(w)))
        else:
            _check(None, lib().tfhe_b200_keygen_ksk(seed, s0p, s1p, ptr(w)))
        return KeySwitchingKey(w)

    def get(self, i, l, t):      # tlwe.rs:281-283 : get(i,l,t) = KS[i][l][t-1]
        off = ((i * K.KS_T + l) * 3 + (t - 1)) * (K.n + 1)
        return self.words[off:off + K.n + 1]


class BootstrappingKey:
    """BootstrappingKey::new(s_key_tlwe, s_key) (tfhe.rs:119-126); flat torus-domain [n][2l][2][N]."""

    def __init__(self, words):
        self.words = np.ascontiguousarray(words, np.uint32).reshape(-1)
        assert self.words.size == K.BK_WORDS

    @staticmethod
    def new(s_key_tlwelv0, s_key_tlwelv1, seed=None):
        """seed=None: fresh CSPRNG key per call; integer seed: deterministic test generator (INSECURE)."""
        w = np.zeros(K.BK_WORDS, np.uint32)
        s0p, s1p = ptr(np.ascontiguousarray(s_key_tlwelv0, np.uint8)), ptr(np.ascontiguousarray(s_key_tlwelv1, np.uint8))
        if seed is None:
            _check(None, lib().tfhe_b200_keygen_bk_csprng(None, s0p, s1p, ptr(w)))
        else:
            _check(None, lib().tfhe_b200_keygen_bk(seed, s0p, s1p, ptr(w)))
        return BootstrappingKey(w)


class Cryptor:
    """Cryptor::{encrypto, decrypto} for the TLWE strategy on bits (digest.rs:14-33; tlwe.rs:197-241)."""

    @staticmethod
    def encrypto(strategy, s_key, item, seed=None, ct_index0=None):
        """seed=None (default): every call draws a fresh 256-bit ChaCha20 key from getrandom(2) -- masks and noise never repeat.
        (seed, ct_index0) integers select the DETERMINISTIC TEST generator: reproducible, NOT secure, and the caller owns the
        uniqueness of the pair (equal pairs repeat masks and noise)."""
        assert strategy is TLWE
        bits = np.ascontiguousarray(np.atleast_1d(item), np.uint8)
        out = np.zeros((len(bits), K.n + 1), np.uint32)
        if seed is None:
            if ct_index0 is not None:
                raise ValueError("ct_index0 belongs to the deterministic test generator: pass a seed with it")
            _check(None, lib().tfhe_b200_encrypt_bits_csprng(None, ptr(np.ascontiguousarray(s_key, np.uint8)), ptr(bits), len(bits), ptr(out)))
        else:
            if ct_index0 is None:
                raise ValueError("the deterministic test generator needs an explicit ct_index0 (equal (seed, index) pairs repeat noise)")
            _check(None, lib().tfhe_b200_encrypt_bits(seed, ct_index0, ptr(np.ascontiguousarray(s_key, np.uint8)), ptr(bits),
                                                      len(bits), ptr(out)))
        return out

    @staticmethod
    def decrypto(strategy, s_key, rep):
        assert strategy is TLWE
        rep = _u32_batch(rep, K.n + 1)
        bits = np.zeros(len(rep), np.uint8)
        _check(None, lib().tfhe_b200_decrypt_bits(ptr(np.ascontiguousarray(s_key, np.uint8)), ptr(rep), len(rep), ptr(bits)))
        return bits

    @staticmethod
    def phase(s_key, rep):
        rep = _u32_batch(rep, K.n + 1)
        ph = np.zeros(len(rep), np.uint32)
        _check(None, lib().tfhe_b200_phase(ptr(np.ascontiguousarray(s_key, np.uint8)), ptr(rep), len(rep), ptr(ph)))
        return ph


# ---- flat file format (the reference has no serialisation) ----
_FILE_SHAPES = {K.FILE_SECRET: (np.uint8, (K.n + K.N,)), K.FILE_BK: (np.uint32, (K.BK_WORDS,)), K.FILE_KSK: (np.uint32, (K.KSK_WORDS,)),
                K.FILE_TLWE0: (np.uint32, (K.n + 1,)), K.FILE_TLWE1: (np.uint32, (K.N + 1,)), K.FILE_TRLWE: (np.uint32, (2, K.N)),
                K.FILE_TRGSW: (np.uint32, (2 * K.L, 2, K.N))}


def _file_check(rc):
    if rc != K.OK:
        msg = lib().tfhe_b200_file_last_error()
        raise TfheError(rc, msg.decode() if msg else "")


def save(path, kind, array):
    """Write keys / ciphertext batches in the flat little-endian format of rustfhe_b200/csrc/wire.cpp."""
    dtype, shape = _FILE_SHAPES[kind]
    a = np.ascontiguousarray(array, dtype)
    rec = int(np.prod(shape))
    if a.size % rec:
        raise ValueError(f"array of {a.size} elements is not a whole number of records of {rec}")
    _file_check(lib().tfhe_b200_file_write(str(path).encode(), kind, ptr(a.reshape(-1)), a.size // rec))


def load(path, kind=None):
    """Read a file written by `save`; returns (kind, array[count, *record shape]) (keys: the single record)."""
    k, cnt, nbytes = C.c_int(), C.c_uint64(), C.c_uint64()
    _file_check(lib().tfhe_b200_file_info(str(path).encode(), C.byref(k), C.byref(cnt), C.byref(nbytes)))
    if kind is not None and k.value != kind:
        raise TfheError(K.ERR_IO, f"{path}: holds kind {k.value}, expected {kind}")
    dtype, shape = _FILE_SHAPES[k.value]
    out = np.empty(nbytes.value // np.dtype(dtype).itemsize, dtype)
    _file_check(lib().tfhe_b200_file_read(str(path).encode(), k.value, ptr(out), nbytes.value))
    out = out.reshape((cnt.value,) + shape)
    if k.value in (K.FILE_SECRET, K.FILE_BK, K.FILE_KSK):
        out = out[0]
    return k.value, out


class TRLWERep:
    """Level-1 ring ciphertexts [B][2][N] (trlwe.rs:13-16): the operations the reference exposes below the gate level."""

    @staticmethod
    def sample_extract_index(engine, rep, index):   # trlwe.rs:110-121
        return engine.sample_extract_batch(rep, index)


class TRGSWRep:
    """Torus-domain TRGSW samples [ntrgsw][2l][2][N] (trgsw.rs:23-26)."""

    @staticmethod
    def cross(engine, trgsw, rhs):                  # trgsw.rs:264-314
        return engine.external_product_batch(trgsw, rhs)

    @staticmethod
    def cmux(engine, trgsw, rep_1, rep_0):          # trgsw.rs:315-330
        return engine.cmux_batch(trgsw, rep_1, rep_0)


class DeviceGroup:
    """Owner of one tfhe_b200_group: ONE process driving several GPUs of a box through the C ABI (rustfhe_b200/csrc/group.cu).
    Keys are replicated once by an NCCL broadcast inside the library; a batch is cut into contiguous shards, one per GPU."""

    def __init__(self, devices=None, decomp_mask=K.MASK_FAITHFUL):
        self._l = lib()
        self._g = C.c_void_p()
        prm = K.Params()
        self._l.tfhe_b200_default_params(C.byref(prm))
        prm.decomp_mask = decomp_mask
        if devices is None:
            arr, n = None, 0
        else:
            n = len(devices)
            arr = (C.c_int * n)(*devices)
        rc = self._l.tfhe_b200_group_create(C.byref(prm), arr, n, C.byref(self._g))
        if rc != K.OK:
            self._g = None
            msg = self._l.tfhe_b200_group_last_error(None)
            raise TfheError(rc, msg.decode() if msg else "")
        self.size = self._l.tfhe_b200_group_size(self._g)

    def close(self):
        if getattr(self, "_g", None):
            self._l.tfhe_b200_group_destroy(self._g)
            self._g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != K.OK:
            msg = self._l.tfhe_b200_group_last_error(self._g)
            raise TfheError(rc, msg.decode() if msg else "")

    def load_bk(self, bk_words):
        self._ck(self._l.tfhe_b200_group_load_bk(self._g, ptr(np.ascontiguousarray(bk_words, np.uint32).reshape(-1))))

    def load_ksk(self, ksk_words):
        self._ck(self._l.tfhe_b200_group_load_ksk(self._g, ptr(np.ascontiguousarray(ksk_words, np.uint32).reshape(-1))))

    def keygen(self, seed, s_key_tlwelv0, s_key_tlwelv1):
        """seed=None: ChaCha20 keyed from getrandom(2); integer: deterministic test generator (INSECURE)."""
        s0p, s1p = ptr(np.ascontiguousarray(s_key_tlwelv0, np.uint8)), ptr(np.ascontiguousarray(s_key_tlwelv1, np.uint8))
        if seed is None:
            self._ck(self._l.tfhe_b200_group_keygen_csprng(self._g, None, s0p, s1p))
        else:
            self._ck(self._l.tfhe_b200_group_keygen(self._g, seed, s0p, s1p))

    def reserve(self, max_batch):
        self._ck(self._l.tfhe_b200_group_reserve(self._g, max_batch))

    def shard(self, B, rank):
        first, count = C.c_size_t(), C.c_size_t()
        self._l.tfhe_b200_group_shard(self._g, B, rank, C.byref(first), C.byref(count))
        return first.value, count.value

    def gate_batch(self, op, in0, in1=None):
        in0 = _u32_batch(in0, K.n + 1)
        in1 = None if in1 is None else _u32_batch(in1, K.n + 1)
        out = np.empty_like(in0)
        self._ck(self._l.tfhe_b200_group_gate_batch(self._g, op, ptr(in0), ptr(in1), ptr(out), len(in0)))
        return out

    def gate_batch_async(self, op, in0, in1, out):
        """in0 / in1 / out: C-contiguous uint32 [B][n+1] arrays that stay alive (and should be pinned) until sync()."""
        self._ck(self._l.tfhe_b200_group_gate_batch_async(self._g, op, ptr(in0), ptr(in1), ptr(out), len(in0)))

    def sync(self):
        self._ck(self._l.tfhe_b200_group_sync(self._g))

    def ctx_stats(self, rank):
        st = K.Stats()
        ctx = self._l.tfhe_b200_group_ctx(self._g, rank)
        _check(ctx, self._l.tfhe_b200_get_stats(ctx, C.byref(st)))
        return {f[0]: getattr(st, f[0]) for f in K.Stats._fields_}


class DeviceEngine:
    """Owner of one tfhe_b200_ctx (one CUDA device). Thin, explicit wrapper of the C ABI."""

    def __init__(self, device=0, decomp_mask=K.MASK_FAITHFUL):
        self._l = lib()
        self._ctx = C.c_void_p()
        prm = K.Params()
        self._l.tfhe_b200_default_params(C.byref(prm))
        prm.decomp_mask = decomp_mask
        rc = self._l.tfhe_b200_ctx_create(C.byref(prm), device, C.byref(self._ctx))
        if rc != K.OK:
            self._ctx = None
            _check(None, rc)
        self.device = device

    def close(self):
        if getattr(self, "_ctx", None):
            self._l.tfhe_b200_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _check(self._ctx, rc)

    # keys
    def load_bk(self, bk_words):
        self._ck(self._l.tfhe_b200_load_bk(self._ctx, ptr(np.ascontiguousarray(bk_words, np.uint32).reshape(-1))))

    def load_ksk(self, ksk_words):
        self._ck(self._l.tfhe_b200_load_ksk(self._ctx, ptr(np.ascontiguousarray(ksk_words, np.uint32).reshape(-1))))

    def load_bk_device(self, dev_ptr, stream=0):
        self._ck(self._l.tfhe_b200_load_bk_device(self._ctx, C.c_void_p(dev_ptr), C.c_void_p(stream)))

    def load_ksk_device(self, dev_ptr, stream=0):
        self._ck(self._l.tfhe_b200_load_ksk_device(self._ctx, C.c_void_p(dev_ptr), C.c_void_p(stream)))

    def keygen_device(self, seed, s_key_tlwelv0, s_key_tlwelv1):
        """BootstrappingKey::new + KeySwitchingKey::new on the device.  seed=None: ChaCha20 keyed from getrandom(2); 32 bytes: that
        ChaCha20 key (bit-identical to the host *_csprng keygen with the same key); integer: deterministic test generator
        (INSECURE; bit-identical to the host keygen with this seed)."""
        s0p, s1p = ptr(np.ascontiguousarray(s_key_tlwelv0, np.uint8)), ptr(np.ascontiguousarray(s_key_tlwelv1, np.uint8))
        if seed is None or isinstance(seed, (bytes, bytearray)):
            key = None if seed is None else (C.c_uint8 * 32).from_buffer_copy(bytes(seed))
            self._ck(self._l.tfhe_b200_keygen_device_csprng(self._ctx, key, s0p, s1p))
        else:
            self._ck(self._l.tfhe_b200_keygen_device(self._ctx, seed, s0p, s1p))

    def export_bk(self):
        w = np.empty(K.BK_WORDS, np.uint32)
        self._ck(self._l.tfhe_b200_export_bk(self._ctx, ptr(w)))
        return BootstrappingKey(w)

    def export_ksk(self):
        w = np.empty(K.KSK_WORDS, np.uint32)
        self._ck(self._l.tfhe_b200_export_ksk(self._ctx, ptr(w)))
        return KeySwitchingKey(w)

    def export_bk_device(self, dev_ptr, stream=0):
        self._ck(self._l.tfhe_b200_export_bk_device(self._ctx, C.c_void_p(dev_ptr), C.c_void_p(stream)))

    def export_ksk_device(self, dev_ptr, stream=0):
        self._ck(self._l.tfhe_b200_export_ksk_device(self._ctx, C.c_void_p(dev_ptr), C.c_void_p(stream)))

    def encrypt_bits_device(self, seed, ct_index0, s_key, bits_ptr, B, out_ptr, stream=0):
        """seed=None: fresh ChaCha20 key from getrandom(2) (ct_index0 ignored); integer seed: deterministic test generator."""
        if seed is None:
            self._ck(self._l.tfhe_b200_encrypt_bits_device_csprng(self._ctx, None, ptr(np.ascontiguousarray(s_key, np.uint8)),
                                                                  C.c_void_p(bits_ptr), B, C.c_void_p(out_ptr), C.c_void_p(stream)))
            return
        self._ck(self._l.tfhe_b200_encrypt_bits_device(self._ctx, seed, ct_index0, ptr(np.ascontiguousarray(s_key, np.uint8)),
                                                       C.c_void_p(bits_ptr), B, C.c_void_p(out_ptr), C.c_void_p(stream)))

    def decrypt_bits_device(self, s_key, ct_ptr, B, bits_ptr=None, phase_ptr=None, stream=0):
        self._ck(self._l.tfhe_b200_decrypt_bits_device(self._ctx, ptr(np.ascontiguousarray(s_key, np.uint8)), C.c_void_p(ct_ptr), B,
                                                       C.c_void_p(bits_ptr or None), C.c_void_p(phase_ptr or None), C.c_void_p(stream)))

    def set_decomp_mask(self, mask):
        self._ck(self._l.tfhe_b200_set_decomp_mask(self._ctx, mask))

    def set_key_slices(self, slices):
        """1 (default) = FFT64 with exact rounding, 2 = NTT with two 16-bit key slices (both exact for honest keys), 3 = NTT with three
        11-bit slices, exact in the worst case (include/tfhe_b200.h)."""
        self._ck(self._l.tfhe_b200_set_key_slices(self._ctx, slices))

    def set_batch_overlap(self, mode):
        """-1 = decide per call (default), 0 = batches run alone, 1 = several batches are kept in flight (include/tfhe_b200.h)."""
        self._ck(self._l.tfhe_b200_set_batch_overlap(self._ctx, mode))

    def reset_stats(self):
        self._ck(self._l.tfhe_b200_reset_stats(self._ctx))

    def stats(self):
        s = K.Stats()
        self._ck(self._l.tfhe_b200_get_stats(self._ctx, C.byref(s)))
        return {f: getattr(s, f) for f, _ in K.Stats._fields_}

    # hot path, host buffers
    def gate_batch(self, op, in0, in1=None):
        in0 = _u32_batch(in0, K.n + 1)
        if in1 is not None:
            in1 = _u32_batch(in1, K.n + 1)
            assert in1.shape == in0.shape
        out = np.empty_like(in0)
        self._ck(self._l.tfhe_b200_gate_batch(self._ctx, op, ptr(in0), ptr(in1), ptr(out), len(in0)))
        return out

    def gate_batch_mixed(self, ops, in0, in1):
        """One launch for gates with different opcodes (`ops`: uint8 [B]); in1 rows of NOT / COPY gates are ignored."""
        ops = np.ascontiguousarray(ops, np.uint8)
        in0 = _u32_batch(in0, K.n + 1)
        in1 = _u32_batch(in1, K.n + 1)
        assert len(ops) == len(in0) == len(in1)
        out = np.empty_like(in0)
        self._ck(self._l.tfhe_b200_gate_batch_mixed(self._ctx, ptr(ops), ptr(in0), ptr(in1), ptr(out), len(in0)))
        return out

    def gate_batch_async(self, op, in0, in1, out):
        """Asynchronous form: `in0`, `in1`, `out` are caller-owned (pinned) uint32 arrays [B][n+1] that must stay
        alive until `sync()`; consecutive calls overlap on the device."""
        assert in0.dtype == np.uint32 and out.dtype == np.uint32 and in0.shape == out.shape and in0.shape[1] == K.n + 1
        self._ck(self._l.tfhe_b200_gate_batch_async(self._ctx, op, ptr(in0), ptr(in1), ptr(out), len(in0)))

    def sync(self):
        self._ck(self._l.tfhe_b200_sync(self._ctx))

    def reserve(self, max_batch):
        """Pre-allocate the internal workspaces for batches of up to `max_batch` gates."""
        self._ck(self._l.tfhe_b200_reserve(self._ctx, max_batch))

    def mux_batch(self, control, in0, in1):
        control, in0, in1 = (_u32_batch(x, K.n + 1) for x in (control, in0, in1))
        out = np.empty_like(in0)
        self._ck(self._l.tfhe_b200_mux_batch(self._ctx, ptr(control), ptr(in0), ptr(in1), ptr(out), len(in0)))
        return out

    # hot path, device pointers (ints) on a given cudaStream_t (int)
    def gate_batch_device(self, op, in0_ptr, in1_ptr, out_ptr, B, stream=0):
        self._ck(self._l.tfhe_b200_gate_batch_device(self._ctx, op, C.c_void_p(in0_ptr), C.c_void_p(in1_ptr or None),
                                                     C.c_void_p(out_ptr), B, C.c_void_p(stream)))

    def mux_batch_device(self, c_ptr, in0_ptr, in1_ptr, out_ptr, B, stream=0):
        self._ck(self._l.tfhe_b200_mux_batch_device(self._ctx, C.c_void_p(c_ptr), C.c_void_p(in0_ptr), C.c_void_p(in1_ptr),
                                                    C.c_void_p(out_ptr), B, C.c_void_p(stream)))

    # step-level entries
    def blind_rotate_batch(self, lin, nsteps=K.n):
        lin = _u32_batch(lin, K.n + 1)
        out = np.empty((len(lin), 2, K.N), np.uint32)
        self._ck(self._l.tfhe_b200_blind_rotate_batch(self._ctx, ptr(lin), nsteps, ptr(out), len(lin)))
        return out

    def bootstrap_lv1_batch(self, lin):
        lin = _u32_batch(lin, K.n + 1)
        out = np.empty((len(lin), K.N + 1), np.uint32)
        self._ck(self._l.tfhe_b200_bootstrap_lv1_batch(self._ctx, ptr(lin), ptr(out), len(lin)))
        return out

    def keyswitch_batch(self, lwe1):
        lwe1 = _u32_batch(lwe1, K.N + 1)
        out = np.empty((len(lwe1), K.n + 1), np.uint32)
        self._ck(self._l.tfhe_b200_keyswitch_batch(self._ctx, ptr(lwe1), ptr(out), len(lwe1)))
        return out

    def external_product_batch(self, trgsw, trlwe):
        trgsw = np.ascontiguousarray(trgsw, np.uint32).reshape(-1, 2 * K.L, 2, K.N)
        trlwe = np.ascontiguousarray(trlwe, np.uint32).reshape(-1, 2, K.N)
        out = np.empty_like(trlwe)
        self._ck(self._l.tfhe_b200_external_product_batch(self._ctx, ptr(trgsw), len(trgsw), ptr(trlwe), ptr(out), len(trlwe)))
        return out

    def external_product_batch_device(self, trgsw_ptr, ntrgsw, trlwe_ptr, out_ptr, B, stream=0):
        self._ck(self._l.tfhe_b200_external_product_batch_device(self._ctx, C.c_void_p(trgsw_ptr), ntrgsw, C.c_void_p(trlwe_ptr),
                                                                 C.c_void_p(out_ptr), B, C.c_void_p(stream)))

    def negacyclic_mul_batch_device(self, a_ptr, d_ptr, out_ptr, B, stream=0):
        self._ck(self._l.tfhe_b200_negacyclic_mul_batch_device(self._ctx, C.c_void_p(a_ptr), C.c_void_p(d_ptr), C.c_void_p(out_ptr), B,
                                                               C.c_void_p(stream)))

    def cmux_batch(self, trgsw, rep_1, rep_0):
        trgsw = np.ascontiguousarray(trgsw, np.uint32).reshape(-1, 2 * K.L, 2, K.N)
        rep_1 = np.ascontiguousarray(rep_1, np.uint32).reshape(-1, 2, K.N)
        rep_0 = np.ascontiguousarray(rep_0, np.uint32).reshape(-1, 2, K.N)
        assert rep_1.shape == rep_0.shape
        out = np.empty_like(rep_1)
        self._ck(self._l.tfhe_b200_cmux_batch(self._ctx, ptr(trgsw), len(trgsw), ptr(rep_1), ptr(rep_0), ptr(out), len(rep_1)))
        return out

    def sample_extract_batch(self, trlwe, index=0):
        trlwe = np.ascontiguousarray(trlwe, np.uint32).reshape(-1, 2, K.N)
        out = np.empty((len(trlwe), K.N + 1), np.uint32)
        self._ck(self._l.tfhe_b200_sample_extract_batch(self._ctx, ptr(trlwe), index, ptr(out), len(trlwe)))
        return out

    def negacyclic_mul_batch(self, a, d):
        a = np.ascontiguousarray(a, np.uint32).reshape(-1, K.N)
        d = np.ascontiguousarray(d, np.int32).reshape(-1, K.N)
        assert a.shape == d.shape
        out = np.empty_like(a)
        self._ck(self._l.tfhe_b200_negacyclic_mul_batch(self._ctx, ptr(a), ptr(d), ptr(out), len(a)))
        return out


class TFHE:
    """`TFHE<TLWE_N, TRLWE_N>`: owns the bootstrapping and key-switching keys and evaluates bootstrapped gates
    (hom_nand/src/tfhe.rs:9-113). Gates take and return TLWERep batches [B][n+1]."""

    def __init__(self, bk, ksk, device=0, decomp_mask=K.MASK_FAITHFUL, engine=None):
        self.bk, self.ksk = bk, ksk
        if engine is not None:           # keys already on the device (TFHE.new_on_device)
            self.engine = engine
            return
        self.engine = DeviceEngine(device, decomp_mask)
        self.engine.load_ksk(ksk.words)
        self.engine.load_bk(bk.words)

    @staticmethod
    def new_on_device(s_key_tlwelv0, s_key_tlwelv1, seed=None, device=0, decomp_mask=K.MASK_FAITHFUL):
        """TFHE::new with both keys generated ON the device (same seed -> same keys as `TFHE.new`); `bk` / `ksk` stay None
        until exported with `engine.export_bk()` / `engine.export_ksk()`."""
        eng = DeviceEngine(device, decomp_mask)
        eng.keygen_device(seed, s_key_tlwelv0, s_key_tlwelv1)
        return TFHE(None, None, device, decomp_mask, engine=eng)

    @staticmethod
    def new(s_key_tlwelv0, s_key_tlwelv1, seed=None, device=0, decomp_mask=K.MASK_FAITHFUL):
        """TFHE::new (tfhe.rs:21-25): ksk = KeySwitchingKey::new(lv1, &lv0); bk = BootstrappingKey::new(lv0, &lv1)."""
        ksk = KeySwitchingKey.new(s_key_tlwelv1, s_key_tlwelv0, seed)
        bk = BootstrappingKey.new(s_key_tlwelv0, s_key_tlwelv1, seed)
        return TFHE(bk, ksk, device, decomp_mask)

    def hom_nand(self, input_0, input_1):
        return self.engine.gate_batch(K.NAND, input_0, input_1)

    def hom_and(self, input_0, input_1):
        return self.engine.gate_batch(K.AND, input_0, input_1)

    def hom_or(self, input_0, input_1):
        return self.engine.gate_batch(K.OR, input_0, input_1)

    def hom_xor(self, input_0, input_1):
        return self.engine.gate_batch(K.XOR, input_0, input_1)

    def hom_not(self, input):
        return self.engine.gate_batch(K.NOT, input)

    def hom_mux(self, control, input_0, input_1):
        """(input_1 & control) | (input_0 & !control) (tfhe.rs:27-40)."""
        return self.engine.mux_batch(control, input_0, input_1)

    def bootstrap(self, tlwelv0):
        return self.engine.gate_batch(K.COPY, tlwelv0)

    def close(self):
        self.engine.close()
