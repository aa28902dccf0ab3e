#!/usr/bin/env python3
"""Stress of the 2-SM cluster latency kernel's DSMEM exchange (GPU box only): the same random gate batches through the cluster
shape (K5FL3, default for batches of at most #SMs/2 gates) and through the one-SM-per-gate shape (K5FL2,
TFHE_B200_F64_CLUSTER=0) of the library in TFHE_B200_LIB; every output ciphertext must be the same bits.
With TFHE_B200_KEY_SLICES=2/3 in the environment: the NTT cluster kernel against the NTT one-CTA kernel (TFHE_B200_BR_VARIANT=9).
usage: python tools/gpu/pair_stress.py [rounds] [B]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import rustfhe_b200 as R
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 48
    seed = 0x5EED0001
    sk = R.SecretKeys.generate(seed)
    os.environ.pop("TFHE_B200_BR_VARIANT", None)
    os.environ.pop("TFHE_B200_SLAB_TMA", None)
    os.environ.pop("TFHE_B200_F64_CLUSTER", None)
    pair = R.TFHE.new_on_device(sk.s_key_tlwelv0, sk.s_key_tlwelv1, seed).engine
    if os.environ.get("TFHE_B200_KEY_SLICES") in ("2", "3"):
        os.environ["TFHE_B200_BR_VARIANT"] = "9"
        os.environ["TFHE_B200_SLAB_TMA"] = "0"     # NTT modes: the plain one-CTA-per-gate kernel (keys streamed from L2)
    else:
        os.environ["TFHE_B200_F64_CLUSTER"] = "0"  # FFT64 mode: one SM per gate is the reference shape
    solo = R.TFHE.new_on_device(sk.s_key_tlwelv0, sk.s_key_tlwelv1, seed).engine
    rng = np.random.default_rng(7)
    bad = 0
    for r in range(rounds):
        b = B if r % 2 == 0 else int(rng.integers(1, B + 1))
        x = rng.integers(0, 2, b).astype(np.uint8)
        y = rng.integers(0, 2, b).astype(np.uint8)
        cx = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, x, seed=seed + 100 + 2 * r, ct_index0=0)
        cy = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, y, seed=seed + 101 + 2 * r, ct_index0=0)
        op = [R.NAND, R.AND, R.OR, R.XOR][r % 4]
        o1 = pair.gate_batch(op, cx, cy)
        o2 = solo.gate_batch(op, cx, cy)
        diff = int((o1 != o2).any(axis=1).sum())
        bad += diff
        if diff:
            print(f"round {r}: {diff} of {b} gates differ", flush=True)
    tag = os.path.basename(os.environ.get("TFHE_B200_LIB", "default"))
    print(f"{tag}: {rounds} rounds of up to {B} gates, gates with differing ciphertexts: {bad}", flush=True)
    pair.close(); solo.close()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
