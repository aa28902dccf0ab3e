"""Small end-to-end run for compute-sanitizer: latency shape (cluster pair), throughput shape, mux, external product."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rustfhe_b200 as R
seed = 0x5EED0001
sk = R.SecretKeys.generate(seed)
tfhe = R.TFHE.new_on_device(sk.s_key_tlwelv0, sk.s_key_tlwelv1, seed)
rng = np.random.default_rng(5)
for B in (3, 150):
    x, y = rng.integers(0, 2, B).astype(np.uint8), rng.integers(0, 2, B).astype(np.uint8)
    c0 = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, x, seed=1, ct_index0=0)
    c1 = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, y, seed=2, ct_index0=0)
    out = tfhe.hom_nand(c0, c1)
    assert np.array_equal(R.Cryptor.decrypto(R.TLWE, sk.s_key_tlwelv0, out), 1 - (x & y)), B
    print("nand ok", B, flush=True)
c = rng.integers(0, 2, 3).astype(np.uint8)
cc = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, c, seed=3, ct_index0=0)
out = tfhe.hom_mux(cc, c0[:3], c1[:3])
print("mux ok", np.array_equal(R.Cryptor.decrypto(R.TLWE, sk.s_key_tlwelv0, out), np.where(c == 1, y[:3], x[:3])), flush=True)
trl = rng.integers(0, 2 ** 32, (3, 2, 1024), dtype=np.uint64).astype(np.uint32)
trg = rng.integers(0, 2 ** 32, (2, 6, 2, 1024), dtype=np.uint64).astype(np.uint32)
tfhe.engine.external_product_batch(trg, trl)
tfhe.engine.cmux_batch(trg, trl, trl[::-1].copy())
print("extprod/cmux ok", flush=True)
tfhe.close()
