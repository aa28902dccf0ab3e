"""ctypes binding of include/tfhe_b200.h (the C ABI is the product boundary; this file is only glue)."""
import ctypes as C
import os

import numpy as np

from . import build as _build

HERE = os.path.dirname(os.path.abspath(__file__))

OK, ERR_PARAM, ERR_CUDA, ERR_STATE, ERR_NOMEM = 0, 1, 2, 3, 4
NAND, AND, OR, XOR, NOT, COPY, ANDNY = range(7)
MASK_FAITHFUL, MASK_CORRECTED = 0x02084000, 0x02082000
n, N, L, KS_T = 635, 1024, 3, 8
BK_WORDS = n * 2 * L * 2 * N
KSK_WORDS = N * KS_T * 3 * (n + 1)

# every symbol include/tfhe_b200.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "tfhe_b200_default_params", "tfhe_b200_ctx_create", "tfhe_b200_ctx_destroy", "tfhe_b200_last_error",
    "tfhe_b200_set_decomp_mask", "tfhe_b200_set_key_slices", "tfhe_b200_set_batch_overlap", "tfhe_b200_get_stats", "tfhe_b200_reset_stats", "tfhe_b200_load_bk", "tfhe_b200_load_bk_device",
    "tfhe_b200_load_ksk", "tfhe_b200_load_ksk_device", "tfhe_b200_gate_batch", "tfhe_b200_gate_batch_device",
    "tfhe_b200_gate_batch_async", "tfhe_b200_sync", "tfhe_b200_reserve", "tfhe_b200_gate_batch_mixed",
    "tfhe_b200_gate_batch_mixed_device", "tfhe_b200_circuit_create", "tfhe_b200_circuit_run_device", "tfhe_b200_circuit_destroy",
    "tfhe_b200_bootstrap_batch", "tfhe_b200_mux_batch", "tfhe_b200_mux_batch_device", "tfhe_b200_blind_rotate_batch",
    "tfhe_b200_bootstrap_lv1_batch", "tfhe_b200_keyswitch_batch", "tfhe_b200_external_product_batch",
    "tfhe_b200_negacyclic_mul_batch", "tfhe_b200_external_product_batch_device", "tfhe_b200_negacyclic_mul_batch_device", "tfhe_b200_keygen_secret", "tfhe_b200_keygen_bk", "tfhe_b200_keygen_ksk",
    "tfhe_b200_encrypt_bits", "tfhe_b200_phase", "tfhe_b200_decrypt_bits", "tfhe_b200_version",
    "tfhe_b200_keygen_device", "tfhe_b200_export_bk", "tfhe_b200_export_ksk", "tfhe_b200_export_bk_device",
    "tfhe_b200_export_ksk_device", "tfhe_b200_encrypt_bits_device",
    "tfhe_b200_decrypt_bits_device", "tfhe_b200_cmux_batch", "tfhe_b200_sample_extract_batch",
    "tfhe_b200_random_bytes", "tfhe_b200_keygen_secret_csprng", "tfhe_b200_keygen_bk_csprng", "tfhe_b200_keygen_ksk_csprng",
    "tfhe_b200_encrypt_bits_csprng", "tfhe_b200_keygen_device_csprng", "tfhe_b200_encrypt_bits_device_csprng",
    "tfhe_b200_group_create", "tfhe_b200_group_destroy", "tfhe_b200_group_last_error", "tfhe_b200_group_size", "tfhe_b200_group_ctx",
    "tfhe_b200_group_load_bk", "tfhe_b200_group_load_ksk", "tfhe_b200_group_keygen_csprng", "tfhe_b200_group_keygen",
    "tfhe_b200_group_reserve", "tfhe_b200_group_shard", "tfhe_b200_group_gate_batch", "tfhe_b200_group_gate_batch_async",
    "tfhe_b200_group_sync", "tfhe_b200_host_alloc", "tfhe_b200_host_free",
    "tfhe_b200_circuit_shape", "tfhe_b200_circuit_level_gates", "tfhe_b200_circuit_run_level_device", "tfhe_b200_circuit_scatter_level_device",
    "tfhe_b200_group_circuit_create", "tfhe_b200_group_circuit_run", "tfhe_b200_group_circuit_stats", "tfhe_b200_group_circuit_destroy",
    "tfhe_b200_file_write", "tfhe_b200_file_info", "tfhe_b200_file_read", "tfhe_b200_file_last_error",
]
FILE_SECRET, FILE_BK, FILE_KSK, FILE_TLWE0, FILE_TLWE1, FILE_TRLWE, FILE_TRGSW = range(1, 8)
ERR_IO = 5


class Params(C.Structure):
    _fields_ = [("n", C.c_int32), ("N", C.c_int32), ("l", C.c_int32), ("bgbit", C.c_int32), ("ks_t", C.c_int32),
                ("ks_basebit", C.c_int32), ("mu", C.c_uint32), ("decomp_mask", C.c_uint32)]


class Stats(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("last_blind_rotate_ms", C.c_float), ("last_keyswitch_ms", C.c_float),
                ("avg_blind_rotate_ms", C.c_float), ("avg_keyswitch_ms", C.c_float), ("timed_launches", C.c_uint64),
                ("last_batch", C.c_uint64), ("gates_per_cta", C.c_int32), ("sm_count", C.c_int32),
                ("device_key_bytes", C.c_uint64), ("key_slices", C.c_int32), ("reserved", C.c_int32)]


_lib = None


def lib():
    """Load librustfhe_b200.so (building it in-tree when sources are newer). Raises if it cannot be loaded: there is
    no Python/CPU fallback for the compute path."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("TFHE_B200_LIB") or _build.build()  # override: experiment builds of the same sources
    l = C.CDLL(path)
    vp, sz, u64, i32, u32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int, C.c_uint32
    sig = {
        "tfhe_b200_default_params": (i32, [C.POINTER(Params)]),
        "tfhe_b200_ctx_create": (i32, [C.POINTER(Params), i32, C.POINTER(vp)]),
        "tfhe_b200_ctx_destroy": (i32, [vp]),
        "tfhe_b200_last_error": (C.c_char_p, [vp]),
        "tfhe_b200_set_decomp_mask": (i32, [vp, u32]),
        "tfhe_b200_set_key_slices": (i32, [vp, i32]),
        "tfhe_b200_set_batch_overlap": (i32, [vp, i32]),
        "tfhe_b200_get_stats": (i32, [vp, C.POINTER(Stats)]),
        "tfhe_b200_reset_stats": (i32, [vp]),
        "tfhe_b200_load_bk": (i32, [vp, vp]),
        "tfhe_b200_load_bk_device": (i32, [vp, vp, vp]),
        "tfhe_b200_load_ksk": (i32, [vp, vp]),
        "tfhe_b200_load_ksk_device": (i32, [vp, vp, vp]),
        "tfhe_b200_gate_batch": (i32, [vp, i32, vp, vp, vp, sz]),
        "tfhe_b200_gate_batch_device": (i32, [vp, i32, vp, vp, vp, sz, vp]),
        "tfhe_b200_gate_batch_async": (i32, [vp, i32, vp, vp, vp, sz]),
        "tfhe_b200_sync": (i32, [vp]),
        "tfhe_b200_reserve": (i32, [vp, sz]),
        "tfhe_b200_gate_batch_mixed": (i32, [vp, vp, vp, vp, vp, sz]),
        "tfhe_b200_gate_batch_mixed_device": (i32, [vp, vp, vp, vp, vp, sz, vp]),
        "tfhe_b200_circuit_create": (i32, [vp, sz, vp, vp, vp, vp, vp, sz, C.POINTER(vp)]),
        "tfhe_b200_circuit_run_device": (i32, [vp, vp, vp, vp]),
        "tfhe_b200_circuit_destroy": (i32, [vp, vp]),
        "tfhe_b200_circuit_shape": (i32, [vp, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)]),
        "tfhe_b200_circuit_level_gates": (i32, [vp, sz, C.POINTER(sz)]),
        "tfhe_b200_circuit_run_level_device": (i32, [vp, vp, sz, sz, sz, vp, vp, vp]),
        "tfhe_b200_circuit_scatter_level_device": (i32, [vp, vp, sz, vp, vp, vp]),
        "tfhe_b200_group_circuit_create": (i32, [vp, sz, vp, vp, vp, vp, vp, sz, sz, C.POINTER(vp)]),
        "tfhe_b200_group_circuit_run": (i32, [vp, vp, vp, sz, vp, vp, sz, vp, sz, vp]),
        "tfhe_b200_group_circuit_stats": (i32, [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(sz)]),
        "tfhe_b200_group_circuit_destroy": (i32, [vp, vp]),
        "tfhe_b200_bootstrap_batch": (i32, [vp, vp, vp, sz]),
        "tfhe_b200_mux_batch": (i32, [vp, vp, vp, vp, vp, sz]),
        "tfhe_b200_mux_batch_device": (i32, [vp, vp, vp, vp, vp, sz, vp]),
        "tfhe_b200_blind_rotate_batch": (i32, [vp, vp, i32, vp, sz]),
        "tfhe_b200_bootstrap_lv1_batch": (i32, [vp, vp, vp, sz]),
        "tfhe_b200_keyswitch_batch": (i32, [vp, vp, vp, sz]),
        "tfhe_b200_external_product_batch": (i32, [vp, vp, sz, vp, vp, sz]),
        "tfhe_b200_negacyclic_mul_batch": (i32, [vp, vp, vp, vp, sz]),
        "tfhe_b200_external_product_batch_device": (i32, [vp, vp, sz, vp, vp, sz, vp]),
        "tfhe_b200_negacyclic_mul_batch_device": (i32, [vp, vp, vp, vp, sz, vp]),
        "tfhe_b200_group_create": (i32, [C.POINTER(Params), vp, i32, C.POINTER(vp)]),
        "tfhe_b200_group_destroy": (i32, [vp]),
        "tfhe_b200_group_last_error": (C.c_char_p, [vp]),
        "tfhe_b200_group_size": (i32, [vp]),
        "tfhe_b200_group_ctx": (vp, [vp, i32]),
        "tfhe_b200_group_load_bk": (i32, [vp, vp]),
        "tfhe_b200_group_load_ksk": (i32, [vp, vp]),
        "tfhe_b200_group_keygen_csprng": (i32, [vp, vp, vp, vp]),
        "tfhe_b200_group_keygen": (i32, [vp, u64, vp, vp]),
        "tfhe_b200_group_reserve": (i32, [vp, sz]),
        "tfhe_b200_group_shard": (None, [vp, sz, i32, C.POINTER(sz), C.POINTER(sz)]),
        "tfhe_b200_group_gate_batch": (i32, [vp, i32, vp, vp, vp, sz]),
        "tfhe_b200_group_gate_batch_async": (i32, [vp, i32, vp, vp, vp, sz]),
        "tfhe_b200_group_sync": (i32, [vp]),
        "tfhe_b200_host_alloc": (i32, [C.POINTER(vp), sz]),
        "tfhe_b200_host_free": (i32, [vp]),
        "tfhe_b200_random_bytes": (i32, [vp, sz]),
        "tfhe_b200_keygen_secret_csprng": (i32, [vp, vp, vp]),
        "tfhe_b200_keygen_bk_csprng": (i32, [vp, vp, vp, vp]),
        "tfhe_b200_keygen_ksk_csprng": (i32, [vp, vp, vp, vp]),
        "tfhe_b200_encrypt_bits_csprng": (i32, [vp, vp, vp, sz, vp]),
        "tfhe_b200_keygen_device_csprng": (i32, [vp, vp, vp, vp]),
        "tfhe_b200_encrypt_bits_device_csprng": (i32, [vp, vp, vp, vp, sz, vp, vp]),
        "tfhe_b200_keygen_secret": (i32, [u64, vp, vp]),
        "tfhe_b200_keygen_bk": (i32, [u64, vp, vp, vp]),
        "tfhe_b200_keygen_ksk": (i32, [u64, vp, vp, vp]),
        "tfhe_b200_encrypt_bits": (i32, [u64, u64, vp, vp, sz, vp]),
        "tfhe_b200_phase": (i32, [vp, vp, sz, vp]),
        "tfhe_b200_decrypt_bits": (i32, [vp, vp, sz, vp]),
        "tfhe_b200_version": (C.c_char_p, []),
        "tfhe_b200_keygen_device": (i32, [vp, u64, vp, vp]),
        "tfhe_b200_export_bk": (i32, [vp, vp]),
        "tfhe_b200_export_ksk": (i32, [vp, vp]),
        "tfhe_b200_export_bk_device": (i32, [vp, vp, vp]),
        "tfhe_b200_export_ksk_device": (i32, [vp, vp, vp]),
        "tfhe_b200_encrypt_bits_device": (i32, [vp, u64, u64, vp, vp, sz, vp, vp]),
        "tfhe_b200_decrypt_bits_device": (i32, [vp, vp, vp, sz, vp, vp, vp]),
        "tfhe_b200_cmux_batch": (i32, [vp, vp, sz, vp, vp, vp, sz]),
        "tfhe_b200_sample_extract_batch": (i32, [vp, vp, i32, vp, sz]),
        "tfhe_b200_file_write": (i32, [C.c_char_p, i32, vp, u64]),
        "tfhe_b200_file_info": (i32, [C.c_char_p, C.POINTER(i32), C.POINTER(u64), C.POINTER(u64)]),
        "tfhe_b200_file_read": (i32, [C.c_char_p, i32, vp, u64]),
        "tfhe_b200_file_last_error": (C.c_char_p, []),
    }
    for name, (res, args) in sig.items():
        f = getattr(l, name)
        f.restype, f.argtypes = res, args
    _lib = l
    return l


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class TfheError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"tfhe_b200 error {code}: {msg}")
        self.code = code
