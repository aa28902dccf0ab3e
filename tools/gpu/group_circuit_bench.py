#!/usr/bin/env python3
"""GPU box only (BASELINE config 4 on 1 and N GPUs): 32-bit adders on encrypted operands evaluated level by level by
tfhe_b200_group_circuit_run on a group of one device and on all devices of the box -- the single ripple-carry and Kogge-Stone
adders (levels narrower than one wave: replicated, no exchange) and K adders side by side (wide levels: sharded, outputs
exchanged over NCCL).  Prints one JSON line per case.  usage: group_circuit_bench.py [K ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import rustfhe_b200 as R
    from rustfhe_b200 import circuit as Cq
    ks = [int(a) for a in sys.argv[1:]] or [16, 64]
    seed = 0x5EED0001
    sk = R.SecretKeys.generate(seed)
    s0 = sk.s_key_tlwelv0
    ndev = torch.cuda.device_count()
    rng = np.random.default_rng(seed + 2)
    cases = [("ripple_carry_nand_x1", Cq.ripple_carry_adder(32), 1), ("kogge_stone_x1", Cq.prefix_adder(32), 1)]
    cases += [(f"kogge_stone_x{k}", Cq.side_by_side(Cq.prefix_adder(32), k), k) for k in ks]
    groups = [[0]] + ([list(range(ndev))] if ndev > 1 else [])
    for devs in groups:
        g = R.DeviceGroup(devs)
        g.keygen(seed, sk.s_key_tlwelv0, sk.s_key_tlwelv1)
        for name, nl, k in cases:
            xs = rng.integers(0, 2 ** 32, k, dtype=np.uint64)
            ys = rng.integers(0, 2 ** 32, k, dtype=np.uint64)
            bits = np.concatenate([np.array([(int(x) >> i) & 1 for i in range(32)] + [(int(y) >> i) & 1 for i in range(32)], np.uint8)
                                   for x, y in zip(xs, ys)])
            cts = R.Cryptor.encrypto(R.TLWE, s0, bits, seed=seed + 300, ct_index0=0)
            gc = Cq.GroupCircuit(g, nl)
            gc.run(cts)
            best = 1e30
            for _ in range(3):
                t0 = time.perf_counter()
                out = gc.run(cts)
                best = min(best, time.perf_counter() - t0)
            got = R.Cryptor.decrypto(R.TLWE, s0, out).reshape(k, 33)
            ok = all(sum(int(b) << i for i, b in enumerate(row)) == int(x) + int(y) for row, x, y in zip(got, xs, ys))
            print(json.dumps({"case": name, "n_gpus": len(devs), "seconds": best, "gates": gc.gates, "levels": gc.levels,
                              "widest_level": max(gc.sizes), "gates_per_s": gc.gates / best, "correct": bool(ok), **gc.last}), flush=True)
            gc.close()
        g.close()


if __name__ == "__main__":
    main()
