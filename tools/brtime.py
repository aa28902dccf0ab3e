#!/usr/bin/env python3
"""Quick kernel timing for experiment builds (GPU box only): gates/s of NAND batches, device resident.

  TFHE_B200_LIB=rustfhe_b200/exp/lib_X.so [TFHE_B200_BR_VARIANT=v] python tools/brtime.py [B ...]
Prints one line per batch size: gates, ms (best of 3), gates/s, blind-rotate / key-switch ms, wrong bits in a sample.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import rustfhe_b200 as R
    from rustfhe_b200 import _capi as K
    sizes = [int(a) for a in sys.argv[1:]] or [1024, 7104]
    seed = 0x5EED0001
    sk = R.SecretKeys.generate(seed)
    tfhe = R.TFHE.new(sk.s_key_tlwelv0, sk.s_key_tlwelv1, seed)
    eng = tfhe.engine
    rng = np.random.default_rng(seed + 3)
    base = 1024
    bx = rng.integers(0, 2, base).astype(np.uint8)
    by = rng.integers(0, 2, base).astype(np.uint8)
    cx = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, bx, seed=seed + 5, ct_index0=0)
    cy = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, by, seed=seed + 6, ct_index0=0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    tag = os.path.basename(os.environ.get("TFHE_B200_LIB", "default")) + " v" + os.environ.get("TFHE_B200_BR_VARIANT", "-")
    for B in sizes:
        reps = (B + base - 1) // base
        dx = torch.from_numpy(np.tile(cx, (reps, 1))[:B].view(np.int32)).to(dev)
        dy = torch.from_numpy(np.tile(cy, (reps, 1))[:B].view(np.int32)).to(dev)
        do = torch.empty_like(dx)
        best = 1e30
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            eng.reset_stats()
            e0.record(stream)
            eng.gate_batch_device(K.NAND, dx.data_ptr(), dy.data_ptr(), do.data_ptr(), B, stream.cuda_stream)
            e1.record(stream)
            torch.cuda.synchronize()
            if it:
                best = min(best, e0.elapsed_time(e1))
        st = eng.stats()
        n = min(B, 1024)
        got = R.Cryptor.decrypto(R.TLWE, sk.s_key_tlwelv0, do[:n].cpu().numpy().view(np.uint32))
        wrong = int((got != (1 - (np.tile(bx, reps)[:n] & np.tile(by, reps)[:n]))).sum())
        print(f"{tag:34s} B={B:6d} ms={best:8.3f} gates/s={B / best * 1e3:9.0f} br={st['last_blind_rotate_ms']:.3f} "
              f"ks={st['last_keyswitch_ms']:.3f} wrong={wrong}", flush=True)
        del dx, dy, do
    eng.close()


if __name__ == "__main__":
    main()
