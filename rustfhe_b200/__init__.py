"""rustfhe_b200 -- B200-native TFHE gate-bootstrapping engine behind the `hom_nand` crate's API surface.

Host-side mirror of the reference interface for the bootstrapped-gate path (names follow hom_nand/src/{tfhe,tlwe,digest}.rs):

    keys  = SecretKeys.generate(seed)                       # s_key_tlwelv0 / s_key_tlwelv1   (homnand-bench.rs:10-12)
    tfhe  = TFHE.new(keys.s_key_tlwelv0, keys.s_key_tlwelv1) # TFHE::new                        (tfhe.rs:21-25)
    c     = Cryptor.encrypto(TLWE, keys.s_key_tlwelv0, bits) # Cryptor::encrypto(TLWE, &s, b)  (digest.rs:17-23)
    out   = tfhe.hom_nand(c0, c1)                           # TFHE::hom_nand, batched          (tfhe.rs:41-47)
    bits  = Cryptor.decrypto(TLWE, keys.s_key_tlwelv0, out) # Cryptor::decrypto                (digest.rs:25-32)

Ciphertexts are `TLWERep` batches: numpy uint32 arrays [B][n+1] (word 0 = cipher b, words 1.. = p_key a).
All compute goes through the C ABI in include/tfhe_b200.h (librustfhe_b200.so, CUDA sm_100a).  There is no CPU fallback:
without the library or without a B200 the compute calls raise.
"""
from ._capi import (AND, ANDNY, COPY, MASK_CORRECTED, MASK_FAITHFUL, NAND, NOT, OR, XOR, BK_WORDS, KSK_WORDS, TfheError)
from . import circuit
from ._capi import (FILE_SECRET, FILE_BK, FILE_KSK, FILE_TLWE0, FILE_TLWE1, FILE_TRLWE, FILE_TRGSW)
from .api import (TFHE, TLWE, BootstrappingKey, Cryptor, KeySwitchingKey, SecretKeys, TFHEHelper, TLWEHelper, TLWERep,
                  TRGSWHelper, TRLWEHelper, DeviceEngine, DeviceGroup, TRLWERep, TRGSWRep, save, load)

__all__ = ["circuit", "TFHE", "TLWE", "BootstrappingKey", "Cryptor", "KeySwitchingKey", "SecretKeys", "TFHEHelper", "TLWEHelper",
           "TLWERep", "TRGSWHelper", "TRLWEHelper", "DeviceEngine", "DeviceGroup", "TfheError", "NAND", "AND", "OR", "XOR", "NOT", "COPY",
           "ANDNY", "MASK_FAITHFUL", "MASK_CORRECTED", "BK_WORDS", "KSK_WORDS", "TRLWERep", "TRGSWRep", "save", "load",
           "FILE_SECRET", "FILE_BK", "FILE_KSK", "FILE_TLWE0", "FILE_TLWE1", "FILE_TRLWE", "FILE_TRGSW"]
