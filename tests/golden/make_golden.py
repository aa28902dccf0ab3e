#!/usr/bin/env python3
"""Generate tests/golden/gate_vectors.npz.  Run in the authoring container where /root/reference exists: the
"*_ref*" entries are outputs of the reference's OWN FFT library (oracle/_ref, compiled in place from
/root/reference/utils/src/spqlios) under the restated gate glue; the "*_exact*" entries are the exact-integer layer."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

O.build(ref=True)
assert O.ref_init()
K = O.Keys(0x5EED0001)
x = np.array([0, 1, 0, 1, 1, 1, 0, 0], np.uint8)
y = np.array([0, 0, 1, 1, 0, 1, 1, 0], np.uint8)
c0, c1 = K.encrypt(x, 7000), K.encrypt(y, 7100)
ex = O.gate_exact(K, O.NAND, c0, c1)
rf = O.gate_ref(K, O.NAND, c0, c1)
rng = np.random.default_rng(2024)
trlwe = rng.integers(0, 2 ** 32, 2048, dtype=np.uint64).astype(np.uint32)
xe, xr = np.zeros(2048, np.uint32), np.zeros(2048, np.uint32)
O.lib().orc_external_product_exact(K.bk[:12 * 1024], trlwe, O.MASK_FAITHFUL, xe)
O.lib().orc_ref_external_product_torus(K.bk[:12 * 1024], trlwe, O.MASK_FAITHFUL, xr)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "gate_vectors.npz"), seed=np.uint64(K.seed), x=x, y=y,
                    c0_head=c0[:, :8], c1_head=c1[:, :8], nand_exact=ex, nand_exact_phase=K.phase(ex),
                    nand_bits=K.decrypt(rf), nand_ref_phase=K.phase(rf), xp_trlwe=trlwe, xp_exact=xe, xp_ref=xr)
print("wrote gate_vectors.npz; ref bits", K.decrypt(rf), "want", 1 - (x & y))
