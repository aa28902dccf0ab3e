"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): every launch shape of the blind rotation with a SHORT
step count (the sanitizer slows kernels down by two orders of magnitude), the key switch, hom_mux, the external product and
cmux entries, the key transforms and the device key generation.  Every result is still checked against the host-side
expectation that is cheap to compute (blind_rotate_batch with nsteps steps is deterministic: runs twice, same bits).

  compute-sanitizer --tool memcheck  python tools/gpu/sanitize.py
  compute-sanitizer --tool racecheck python tools/gpu/sanitize.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rustfhe_b200 as R  # noqa: E402

NSTEPS = int(os.environ.get("SANITIZE_STEPS", "6"))
seed = 0x5EED0001
sk = R.SecretKeys.generate(seed)
tfhe = R.TFHE.new_on_device(sk.s_key_tlwelv0, sk.s_key_tlwelv1, seed)   # bk_fill / polymul / gadget / transforms / lwe_rows kernels
eng = tfhe.engine
rng = np.random.default_rng(5)
sm = eng.stats()["sm_count"]


def shapes(tag):
    # (batch, what it exercises): cluster pair, one gate per CTA with bulk-copied slabs, throughput kernel (ragged dealing)
    for B in (3, sm // 2 + 5, sm + 3, 2 * sm + 5, 6 * sm + 7):
        x = rng.integers(0, 2, B).astype(np.uint8)
        c = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, x, seed=11, ct_index0=1000)
        a = eng.blind_rotate_batch(c, NSTEPS)
        b = eng.blind_rotate_batch(c, NSTEPS)
        assert np.array_equal(a, b), (tag, B)
        print(f"{tag}: blind rotate x{NSTEPS} steps, {B} gates: deterministic, gates/CTA {eng.stats()['gates_per_cta']}", flush=True)


shapes("two key slices (default)")
eng.set_key_slices(3)
shapes("three key slices")
eng.set_key_slices(2)
# key switch alone (both kernels are reached through the batch size), sample extract, external product, cmux, poly mul
lwe1 = rng.integers(0, 2 ** 32, (40, 1025), dtype=np.uint64).astype(np.uint32)
k1 = eng.keyswitch_batch(lwe1)
assert np.array_equal(k1, eng.keyswitch_batch(lwe1))
trl = rng.integers(0, 2 ** 32, (5, 2, 1024), dtype=np.uint64).astype(np.uint32)
trg = rng.integers(0, 2 ** 32, (2, 6, 2, 1024), dtype=np.uint64).astype(np.uint32)
eng.external_product_batch(trg, trl)
eng.cmux_batch(trg, trl, trl[::-1].copy())
eng.sample_extract_batch(trl, 17)
eng.negacyclic_mul_batch(trl[:, 0], rng.integers(-32, 32, (5, 1024)).astype(np.int32))
print("key switch / external product / cmux / extract / poly mul ok", flush=True)
if os.environ.get("SANITIZE_FULL_GATE", "1") == "1":
    # one complete gate on the smallest shape: the whole pipeline end to end, decrypt checked
    x, y = np.array([0, 1, 1], np.uint8), np.array([1, 1, 0], np.uint8)
    c0 = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, x, seed=1, ct_index0=0)
    c1 = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, y, seed=2, ct_index0=0)
    out = tfhe.hom_nand(c0, c1)
    assert np.array_equal(R.Cryptor.decrypto(R.TLWE, sk.s_key_tlwelv0, out), 1 - (x & y))
    print("full nand gates ok", flush=True)
tfhe.close()
print("sanitize workload done", flush=True)
