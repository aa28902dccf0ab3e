#!/usr/bin/env python3
"""GPU box only: the FFT64 throughput kernel variant in TFHE_B200_F64_TMEM (argv[1]) against the default kernel -- same bits on
an unevenly dealt batch -- then blind-rotation times of both at a few batch sizes.  usage: f64_variant_check.py VARIANT [B ...]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import rustfhe_b200 as R
    variant = sys.argv[1]
    sizes = [int(a) for a in sys.argv[2:]] or [888, 1184, 1776, 7104]
    seed = 0x5EED0001
    sk = R.SecretKeys.generate(seed)
    engs = {}
    os.environ["TFHE_B200_F64_LATENCY"] = "0"   # every batch above #SMs gates through the throughput kernel
    for flag in ("0", variant):
        os.environ["TFHE_B200_F64_TMEM"] = flag   # 0 = K5F, 1 = K5FT, 2 = K5F2
        engs[flag] = R.TFHE.new_on_device(sk.s_key_tlwelv0, sk.s_key_tlwelv1, seed).engine
    rng = np.random.default_rng(11)
    bad = 0
    for B in (149, 449, 6 * 148 + 5, 2048):
        x = rng.integers(0, 2, B).astype(np.uint8)
        y = rng.integers(0, 2, B).astype(np.uint8)
        cx = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, x, seed=seed + 100 + B, ct_index0=0)
        cy = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, y, seed=seed + 101 + B, ct_index0=0)
        o = {f: e.gate_batch(R.XOR, cx, cy) for f, e in engs.items()}
        diff = int((o["0"] != o[variant]).any(axis=1).sum())
        ok = np.array_equal(R.Cryptor.decrypto(R.TLWE, sk.s_key_tlwelv0, o[variant]), x ^ y)
        print(f"B={B}: gates with differing ciphertexts {diff}, decrypts {'ok' if ok else 'WRONG'}, gates/CTA {engs[variant].stats()['gates_per_cta']}", flush=True)
        bad += diff + (0 if ok else 1)
    for B in sizes:
        x = rng.integers(0, 2, B).astype(np.uint8)
        cx = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, x, seed=seed + 7, ct_index0=0)
        for f, e in engs.items():
            best = 1e30
            for it in range(4):
                e.reset_stats()
                e.gate_batch(R.NAND, cx, cx)
                if it:
                    best = min(best, e.stats()["last_blind_rotate_ms"])
            print(f"variant {f}: B={B} blind rotation {best:.3f} ms = {B / best:.1f} k gates/s in kernel", flush=True)
    for e in engs.values():
        e.close()
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
