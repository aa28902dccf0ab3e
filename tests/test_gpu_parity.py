"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on identical seeded inputs.

Parity ladder:
  P0  kernel vs exact-integer oracle ........ BIT-EXACT (poly mul, external product, blind rotate, extract, key switch,
                                               whole gate ciphertext)
  P1  single external product vs the reference's own FFT ... every coefficient within +-2 ulp (2^-32)
  P2  whole gate vs reference-FFT gate path .. identical decrypted bit, |e_gpu| < 3/32, |e_gpu - e_ref| < 1/16
"""
import numpy as np
import pytest

from conftest import phase_err

pytestmark = pytest.mark.gpu

N, n = 1024, 635
# Default arithmetic: FFT64 (f64 transform with exact rounding) for batches above #SMs gates, the two-slice NTT below and for
# the step-level entry points.  Both are exact on random and real keys (DESIGN.md section 2 has the margins) but by design not on
# the adversarial worst-case KEY below; that vector runs in the worst-case-exact three-slice mode
# (test_exact_mode_three_key_slices, or the whole suite with TFHE_B200_KEY_SLICES=3).
FAST_MODE = __import__("os").environ.get("TFHE_B200_KEY_SLICES") != "3"


def u32(rng, *shape):
    return rng.integers(0, 2 ** 32, size=shape, dtype=np.uint64).astype(np.uint32)


def test_negacyclic_mul_exact(engine, oracle, rng):
    """config 3(i): torus x small-int polynomial, bit-exact vs schoolbook (math.rs:761-843 / 905-952 semantics)."""
    B = 33
    a = u32(rng, B, N)
    d = rng.integers(-32, 32, size=(B, N)).astype(np.int32)
    d[1] = rng.integers(0, 2, size=N)                # binary variant mirroring math.rs:934-951
    a[2], d[2] = 0xFFFFFFFF, 31                      # extreme magnitudes
    a[3], d[3] = 0x80000000, -32
    d[4] = rng.integers(-192, 193, size=N)           # documented bound of the entry
    out = engine.negacyclic_mul_batch(a, d)
    for g in range(B):
        ref = np.zeros(N, np.uint32)
        oracle.lib().orc_negacyclic_mul_schoolbook(a[g], d[g], N, ref)
        assert np.array_equal(out[g], ref), g


def test_negacyclic_mul_large_batch(engine, oracle, rng):
    """config 3(i) at throughput size (above #SMs products: in the default mode the persistent FFT64 kernel polymul_f64_kernel),
    with the extreme vectors of the small test in it, bit-exact vs schoolbook on a sample and identical to the small-batch path."""
    B = 12 * 148 + 7
    a = u32(rng, B, N)
    d = rng.integers(-32, 32, size=(B, N)).astype(np.int32)
    a[2], d[2] = 0xFFFFFFFF, 31
    a[3], d[3] = 0x80000000, -32
    a[5], d[5] = 0x7FFFFFFF, 192                     # the documented bound of the entry, all coefficients extreme
    d[4] = rng.integers(-192, 193, size=N)
    out = engine.negacyclic_mul_batch(a, d)
    for g in np.concatenate([[0, 2, 3, 4, 5, B - 1], rng.choice(B, 4, replace=False)]):
        ref = np.zeros(N, np.uint32)
        oracle.lib().orc_negacyclic_mul_schoolbook(a[g], d[g], N, ref)
        assert np.array_equal(out[g], ref), g
    assert np.array_equal(engine.negacyclic_mul_batch(a[:33], d[:33]), out[:33])


def test_negacyclic_mul_kat_via_device(engine):
    """reference KAT math.rs:761-843: [2,3,4]*[4,5,6] = [-30,-2,43] mod X^3+1 embeds in N=1024 only as a linear product;
    use the N-independent identities X^k * a instead: a * X^1023 * X = -a."""
    a = np.arange(1, N + 1, dtype=np.uint32).reshape(1, N)
    d = np.zeros((1, N), np.int32)
    d[0, 1] = 1                                       # multiply by X == rotate(1) (math.rs:75-84)
    out = engine.negacyclic_mul_batch(a, d)[0]
    assert out[0] == np.uint32(-int(a[0, N - 1]) & 0xFFFFFFFF) and np.array_equal(out[1:], a[0, :-1])


@pytest.mark.parametrize("mask", [0x02084000, 0x02082000])
def test_external_product_exact(engine, oracle, rng, mask):
    """config 3(ii): TRGSW (x) TRLWE, shared and per-item TRGSW, bit-exact vs the integer oracle (trgsw.rs:264-306)."""
    engine.set_decomp_mask(mask)
    try:
        B = 5
        trlwe = u32(rng, B, 2, N)
        trlwe[1] = 0x7DF7C000  # all digits -32
        for ntr in (1, B):
            trgsw = u32(rng, ntr, 6, 2, N)
            if ntr == B and not FAST_MODE:
                trgsw[1] = 0x7FFFFFFF
            out = engine.external_product_batch(trgsw, trlwe)
            for g in range(B):
                ref = np.zeros(2 * N, np.uint32)
                oracle.lib().orc_external_product_exact(trgsw[g % ntr].reshape(-1), trlwe[g].reshape(-1), mask, ref)
                assert np.array_equal(out[g].reshape(-1), ref), (ntr, g)
    finally:
        engine.set_decomp_mask(0x02084000)


def test_external_product_large_batch_shared_trgsw(engine, oracle, rng):
    """config 3(ii) at throughput size: one shared TRGSW against a batch above #SMs products (in the default mode the persistent
    FFT64 kernel external_product_f64_kernel: one product per warp, the TRGSW going round the key ring; ragged last round), plain
    product and cmux form, bit-exact vs the integer oracle on a sample, and identical to the small-batch (NTT) path."""
    B = 8 * 148 * 2 + 3
    trlwe, rep0 = u32(rng, B, 2, N), u32(rng, B, 2, N)
    trlwe[1] = 0x7DF7C000            # all digits -32
    trgsw = u32(rng, 1, 6, 2, N)
    out = engine.external_product_batch(trgsw, trlwe)
    cm = engine.cmux_batch(trgsw, trlwe, rep0)
    idx = np.concatenate([[0, 1, B - 1, B - 2], rng.choice(B, 6, replace=False)])
    for g in idx:
        ref = np.zeros(2 * N, np.uint32)
        oracle.lib().orc_external_product_exact(trgsw[0].reshape(-1), trlwe[g].reshape(-1), 0x02084000, ref)
        assert np.array_equal(out[g].reshape(-1), ref), g
        oracle.lib().orc_external_product_exact(trgsw[0].reshape(-1), (trlwe[g] - rep0[g]).reshape(-1), 0x02084000, ref)
        assert np.array_equal(cm[g].reshape(-1), ref + rep0[g].reshape(-1)), g
    small = engine.external_product_batch(trgsw, trlwe[:40])
    assert np.array_equal(small, out[:40])


def test_external_product_large_batch_per_item_trgsw(engine, oracle, rng):
    """config 3(ii) with one TRGSW per item at throughput size (in the default mode external_product_item_f64_kernel: a warp
    transforms its item's twelve TRGSW polynomials itself; ragged last round), plain product and cmux form, bit-exact vs the
    integer oracle on a sample -- including an all-digits -32 input against a TRGSW of +-2^31 words -- and identical to the
    small-batch path."""
    B = 8 * 148 + 5
    trlwe, rep0 = u32(rng, B, 2, N), u32(rng, B, 2, N)
    trgsw = u32(rng, B, 6, 2, N)
    trlwe[1] = 0x7DF7C000            # all digits -32
    trgsw[1] = 0x80000000            # extreme key words
    out = engine.external_product_batch(trgsw, trlwe)
    cm = engine.cmux_batch(trgsw, trlwe, rep0)
    idx = np.concatenate([[0, 1, B - 1, B - 2], rng.choice(B, 5, replace=False)])
    for g in idx:
        ref = np.zeros(2 * N, np.uint32)
        oracle.lib().orc_external_product_exact(trgsw[g].reshape(-1), trlwe[g].reshape(-1), 0x02084000, ref)
        assert np.array_equal(out[g].reshape(-1), ref), g
        oracle.lib().orc_external_product_exact(trgsw[g].reshape(-1), (trlwe[g] - rep0[g]).reshape(-1), 0x02084000, ref)
        assert np.array_equal(cm[g].reshape(-1), ref + rep0[g].reshape(-1)), g
    # the small-batch path of this mode (two-slice NTT kernel) agrees on every ordinary item; item 1 (all key words 2^31, all
    # digits -32) is outside the bound that path is exact for (mode 3 is the worst-case-exact one, test_exact_mode_three_key_slices)
    small = engine.external_product_batch(trgsw[:40], trlwe[:40])
    keep = np.arange(40) != 1
    assert np.array_equal(small[keep], out[:40][keep])


def test_external_product_vs_reference_fft(engine, oracle, rng):
    """P1: against the reference's own FFT (oracle/_ref): |diff| <= 2 ulp per coefficient (measured -1/0/+1)."""
    if not oracle.ref_init():
        pytest.skip("oracle/_ref/libspqlios_ref.so not built")
    B = 4
    trgsw = u32(rng, 1, 6, 2, N)
    trlwe = u32(rng, B, 2, N)
    out = engine.external_product_batch(trgsw, trlwe)
    worst = 0
    for g in range(B):
        ref = np.zeros(2 * N, np.uint32)
        oracle.lib().orc_ref_external_product_torus(trgsw.reshape(-1), trlwe[g].reshape(-1), 0x02084000, ref)
        diff = (out[g].reshape(-1).astype(np.int64) - ref.astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
        worst = max(worst, int(np.abs(diff).max()))
    assert worst <= 2, worst


@pytest.mark.parametrize("nsteps", [0, 1, 2, 17, 635])
def test_blind_rotate_exact(engine, oracle, keys, rng, nsteps):
    """blind_rotate prefix of nsteps CMUXes, raw TRLWE accumulator bit-exact vs the integer oracle (tfhe.rs:89-113)."""
    B = 3 if nsteps == 635 else 5
    bits = rng.integers(0, 2, B).astype(np.uint8)
    lin = oracle.gate_linear(oracle.NAND, keys.encrypt(bits, 0), keys.encrypt(1 - bits, 1000))
    lin[0, 0] = 0                      # bbar = 0
    if B > 3:
        lin[3, 0] = 0xFFE00000         # bbar = 2047
        lin[4, 0] = 0x80000000         # bbar = 1024
        lin[4, 1] = 0xFFF00000         # abar rounds up to 2048 -> 0
        lin[4, 2] = 0x7FF00000         # abar = 1024
    out = engine.blind_rotate_batch(lin, nsteps)
    for g in range(B):
        ref = np.zeros(2 * N, np.uint32)
        oracle.lib().orc_blind_rotate_exact(keys.exact_handle(), lin[g], 0x02084000, nsteps, ref)
        assert np.array_equal(out[g].reshape(-1), ref), (nsteps, g)


def test_sample_extract_and_keyswitch_exact(engine, oracle, keys, rng):
    """sample_extract_index(0) (trlwe.rs:110-121) and identity_key_switch (tlwe.rs:43-73), bit-exact."""
    B = 19  # not a multiple of the key-switch gate tile
    bits = rng.integers(0, 2, B).astype(np.uint8)
    lin = oracle.gate_linear(oracle.COPY, keys.encrypt(bits, 50))
    lwe1 = engine.bootstrap_lv1_batch(lin)
    trl = engine.blind_rotate_batch(lin, 635)
    for g in range(B):
        ref = np.zeros(N + 1, np.uint32)
        oracle.lib().orc_sample_extract0(trl[g].reshape(-1), ref)
        assert np.array_equal(lwe1[g], ref)
    # key switch on arbitrary level-1 samples, including digit edge cases
    x = u32(rng, B, N + 1)
    x[0, 1:] = 0               # all digits zero
    x[1, 1:] = 0xFFFF8000      # rounding carries out of the top
    x[2, 1:] = 0x55555555
    out = engine.keyswitch_batch(x)
    for g in range(B):
        ref = np.zeros(n + 1, np.uint32)
        oracle.lib().orc_key_switch(keys.ksk, x[g], ref)
        assert np.array_equal(out[g], ref), g
    # and the real thing decrypts
    out = engine.keyswitch_batch(lwe1)
    assert np.array_equal(keys.decrypt(out), bits)


def test_keyswitch_kernel_variants_same_bits(oracle, keys, rng, monkeypatch):
    """The three key-switch kernels (TFHE_B200_KS_VARIANT: 3 = producer / consumer ring, the default; 2 = CTA barrier per stage;
    1 = register tiles) give the same words on ragged batches (1, 17, 33 and 1500 samples: partial tiles, several tiles, every
    split of the key indices), and the oracle's on a sample."""
    import rustfhe_b200 as R
    outs = {}
    xs = {B: u32(rng, B, N + 1) for B in (1, 17, 33, 1500)}
    for v in ("3", "2", "1"):
        monkeypatch.setenv("TFHE_B200_KS_VARIANT", v)
        eng = R.DeviceEngine(0)
        try:
            eng.load_ksk(keys.ksk)
            eng.load_bk(keys.bk)
            outs[v] = {B: eng.keyswitch_batch(x) for B, x in xs.items()}
        finally:
            eng.close()
    for B, x in xs.items():
        assert np.array_equal(outs["3"][B], outs["2"][B]) and np.array_equal(outs["3"][B], outs["1"][B]), B
        for g in (0, B - 1):
            ref = np.zeros(n + 1, np.uint32)
            oracle.lib().orc_key_switch(keys.ksk, x[g], ref)
            assert np.array_equal(outs["3"][B][g], ref), (B, g)


OPS = ["NAND", "AND", "OR", "XOR", "NOT"]


def truth(op, x, y):
    return {"NAND": 1 - (x & y), "AND": x & y, "OR": x | y, "XOR": x ^ y, "NOT": 1 - x}[op]


@pytest.mark.parametrize("op", OPS)
def test_gate_truth_tables_exact_and_reference(engine, oracle, keys, op):
    """config 1 workload (examples/homnand-bench.rs:22-136, tfhe.rs:147-279): truth tables of nand/and/or/xor/not with fresh
    encryptions.  P0: whole output ciphertext bit-exact vs the integer oracle.  P2: vs the reference-FFT gate path."""
    x = np.array([0, 1, 0, 1], np.uint8)
    y = np.array([0, 0, 1, 1], np.uint8)
    if op == "NOT":
        x, y = np.array([0, 1], np.uint8), None
    c0 = keys.encrypt(x, 7000)
    c1 = keys.encrypt(y, 7100) if y is not None else None
    code = getattr(oracle, op)
    out = engine.gate_batch(code, c0, c1)
    want = truth(op, x, y if y is not None else x)
    assert np.array_equal(keys.decrypt(out), want)
    exact = oracle.gate_exact(keys, code, c0, c1)
    assert np.array_equal(out, exact), "ciphertext differs from the exact-integer oracle"
    if oracle.ref_init():
        ref = oracle.gate_ref(keys, code, c0, c1)
        assert np.array_equal(keys.decrypt(ref), want)
        e_gpu, e_ref = phase_err(keys, out, want), phase_err(keys, ref, want)
        assert np.abs(e_gpu).max() < 3 / 32
        assert np.abs(e_gpu - e_ref).max() < 1 / 16


def test_mux(engine, oracle, keys):
    """hom_mux = bootstrap(AND(c,i1) + AND(-c,i0) + 1/8): all 8 input combinations (tfhe.rs:27-40)."""
    c = np.array([0, 0, 0, 0, 1, 1, 1, 1], np.uint8)
    i0 = np.array([0, 0, 1, 1, 0, 0, 1, 1], np.uint8)
    i1 = np.array([0, 1, 0, 1, 0, 1, 0, 1], np.uint8)
    cc, c0, c1 = keys.encrypt(c, 9000), keys.encrypt(i0, 9100), keys.encrypt(i1, 9200)
    out = engine.mux_batch(cc, c0, c1)
    assert np.array_equal(keys.decrypt(out), np.where(c == 1, i1, i0))
    t1 = oracle.gate_exact(keys, oracle.AND, cc, c1)
    t0 = oracle.gate_exact(keys, oracle.ANDNY, cc, c0)
    assert np.array_equal(out, oracle.gate_exact(keys, oracle.OR, t1, t0))


def test_batch_1024_nand(engine, oracle, keys, rng):
    """config 2: 1024 independent NAND gates; all decrypts correct, a 64-gate subset bit-exact vs the oracle and within
    the P2 phase bounds vs the reference-FFT path; phase-noise std within 15% of the reference's."""
    B = 1024
    x = rng.integers(0, 2, B).astype(np.uint8)
    y = rng.integers(0, 2, B).astype(np.uint8)
    c0, c1 = keys.encrypt(x, 100000), keys.encrypt(y, 200000)
    out = engine.gate_batch(oracle.NAND, c0, c1)
    want = 1 - (x & y)
    assert np.array_equal(keys.decrypt(out), want)
    sub = np.arange(0, B, 16)
    assert np.array_equal(out[sub], oracle.gate_exact(keys, oracle.NAND, c0[sub], c1[sub]))
    e_gpu = phase_err(keys, out, want)
    assert np.abs(e_gpu).max() < 3 / 32
    if oracle.ref_init():
        ref = oracle.gate_ref(keys, oracle.NAND, c0[sub], c1[sub])
        e_ref = phase_err(keys, ref, want[sub])
        assert np.abs(e_gpu[sub] - e_ref).max() < 1 / 16
        assert 0.85 < e_gpu.std() / e_ref.std() < 1.15, (e_gpu.std(), e_ref.std())


def test_ragged_and_edge_batches(engine, oracle, keys, rng):
    """empty batch, batch of one (latency path, 1 gate per CTA), odd sizes around the CTA pairing and the SM count."""
    assert engine.gate_batch(oracle.NAND, np.zeros((0, n + 1), np.uint32), np.zeros((0, n + 1), np.uint32)).shape == (0, n + 1)
    for B in (1, 2, 3, 149, 151):
        x = rng.integers(0, 2, B).astype(np.uint8)
        y = rng.integers(0, 2, B).astype(np.uint8)
        out = engine.gate_batch(oracle.XOR, keys.encrypt(x, 300000), keys.encrypt(y, 400000))
        assert np.array_equal(keys.decrypt(out), x ^ y), B


def test_trivial_inputs_like_nander(engine, oracle, keys):
    """nander's leaves are TRIVIAL ciphertexts TLWERep::logic_true/false (tlwe.rs:80-87, nander/src/lib.rs:150-151)."""
    import rustfhe_b200 as R
    t, f = R.TLWERep.logic_true(), R.TLWERep.logic_false()
    for a, b, abit, bbit in ((t, t, 1, 1), (t, f, 1, 0), (f, f, 0, 0)):
        out = engine.gate_batch(oracle.NAND, a, b)
        assert keys.decrypt(out)[0] == 1 - (abit & bbit)
        assert np.array_equal(out, oracle.gate_exact(keys, oracle.NAND, a, b))


def test_error_paths(engine):
    import rustfhe_b200 as R
    with pytest.raises(R.TfheError):
        engine._ck(engine._l.tfhe_b200_gate_batch(engine._ctx, 99, None, None, None, 1))
    fresh = R.DeviceEngine(0)
    try:
        with pytest.raises(R.TfheError) as ei:
            fresh.gate_batch(R.NAND, np.zeros((1, n + 1), np.uint32), np.zeros((1, n + 1), np.uint32))
        assert ei.value.code == 3  # TFHE_B200_ERR_STATE: keys not loaded
    finally:
        fresh.close()


# ---------------------------------------------------------------------------------------------------------------
# device-side key generation / encryption, scheme surface below the gate level
# ---------------------------------------------------------------------------------------------------------------
def test_device_keygen_bit_identical_to_host_and_oracle(oracle, keys):
    """BootstrappingKey::new / KeySwitchingKey::new on the device (tfhe.rs:119-126, tlwe.rs:247-277): same seeded counter
    generator and the same deterministic Gaussian as the host keygen and the oracle keygen -> identical keys, bit for bit."""
    import rustfhe_b200 as R
    eng = R.DeviceEngine(0)
    try:
        eng.keygen_device(keys.seed, keys.s0, keys.s1)
        bk, ksk = eng.export_bk().words, eng.export_ksk().words
        assert np.array_equal(ksk, keys.ksk.reshape(-1))
        assert np.array_equal(bk, keys.bk.reshape(-1))
        # and the device-generated keys evaluate gates: NAND truth table, bit-exact against the oracle with the oracle's keys
        x, y = np.array([0, 0, 1, 1], np.uint8), np.array([0, 1, 0, 1], np.uint8)
        c0, c1 = keys.encrypt(x, 1200), keys.encrypt(y, 1300)
        out = eng.gate_batch(R.NAND, c0, c1)
        assert np.array_equal(keys.decrypt(out), 1 - (x & y))
        assert np.array_equal(out, oracle.gate_exact(keys, oracle.NAND, c0, c1))
    finally:
        eng.close()


def test_device_encrypt_decrypt_bit_identical(engine, oracle, keys, rng):
    """Cryptor::encrypto / decrypto (TLWE, bits) on the device (tlwe.rs:181-240) vs the oracle's encryption, same seed."""
    import torch
    B = 777
    bits = rng.integers(0, 2, B).astype(np.uint8)
    dbits = torch.from_numpy(bits).cuda()
    dct = torch.empty((B, n + 1), dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    engine.encrypt_bits_device(keys.seed, 4242, keys.s0, dbits.data_ptr(), B, dct.data_ptr(), st)
    torch.cuda.synchronize()
    ct = dct.cpu().numpy().view(np.uint32)
    assert np.array_equal(ct, keys.encrypt(bits, 4242, seed=keys.seed))
    dout = torch.empty(B, dtype=torch.uint8, device="cuda")
    dph = torch.empty(B, dtype=torch.int32, device="cuda")
    engine.decrypt_bits_device(keys.s0, dct.data_ptr(), B, dout.data_ptr(), dph.data_ptr(), st)
    torch.cuda.synchronize()
    assert np.array_equal(dout.cpu().numpy(), bits)
    assert np.array_equal(dph.cpu().numpy().view(np.uint32), keys.phase(ct))


@pytest.mark.parametrize("index", [0, 1, 511, 1023])
def test_sample_extract_index(engine, oracle, rng, index):
    """TRLWERep::sample_extract_index(i) for i != 0 too (trlwe.rs:110-121), bit-exact."""
    B = 7
    trl = u32(rng, B, 2, N)
    out = engine.sample_extract_batch(trl, index)
    for g in range(B):
        ref = np.zeros(N + 1, np.uint32)
        oracle.lib().orc_sample_extract(trl[g].reshape(-1), index, ref)
        assert np.array_equal(out[g], ref), (index, g)


def test_cmux_exact_and_selects(engine, oracle, keys, rng):
    """TRGSWRep::cmux(rep_1, rep_0) = cross(rep_1 - rep_0) + rep_0 (trgsw.rs:315-330): bit-exact vs the integer oracle, and
    with real TRGSW(0) / TRGSW(1) samples (two elements of the bootstrapping key) it selects rep_0 / rep_1 up to noise."""
    B = 6
    rep1, rep0 = u32(rng, B, 2, N), u32(rng, B, 2, N)
    trgsw = u32(rng, 2, 6, 2, N)
    out = engine.cmux_batch(trgsw, rep1, rep0)
    for g in range(B):
        ref = np.zeros(2 * N, np.uint32)
        oracle.lib().orc_external_product_exact(trgsw[g % 2].reshape(-1), (rep1[g] - rep0[g]).reshape(-1), 0x02084000, ref)
        assert np.array_equal(out[g].reshape(-1), ref + rep0[g].reshape(-1)), g
    # selection with real TRGSW encryptions of the lv0 key bits: BK_i = TRGSW_{s1}(s0_i)
    i0 = int(np.flatnonzero(keys.s0 == 0)[0])
    i1 = int(np.flatnonzero(keys.s0 == 1)[0])
    bk = keys.bk.reshape(n, 6, 2, N)
    m1 = np.full(N, 0x20000000, np.uint32)
    m0 = np.full(N, 0xE0000000, np.uint32)
    t1 = np.stack([m1, np.zeros(N, np.uint32)])[None]      # trivial TRLWE(+1/8), TRLWE(-1/8)
    t0 = np.stack([m0, np.zeros(N, np.uint32)])[None]
    for idx, want in ((i0, m0), (i1, m1)):
        res = engine.cmux_batch(bk[idx][None], t1, t0)[0]
        ph = np.zeros(N, np.uint32)
        oracle.lib().orc_trlwe_phase(keys.s1, res.reshape(-1), ph)
        err = (ph.astype(np.int64) - want.astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
        assert np.abs(err).max() < 2 ** 32 / 64, idx


@pytest.mark.parametrize("B", [5, 49, 74, 75, 148, 149, 297, 445, 601, 889])
def test_launch_shapes_deterministic_and_exact(engine, oracle, keys, rng, B):
    """Every launch shape of the default mode (one gate per SM on twelve warps -- K5FL2 -- in one, two and three waves up to 3 #SMs
    gates, the one-warp-per-gate throughput kernel K5F with uneven dealing above; with TFHE_B200_KEY_SLICES=2/3 the NTT shapes:
    2-SM cluster per gate, one gate per CTA, K5T / K5) gives the same bits run after run (a shared-memory race would not) and
    matches the exact oracle on a sample; all decrypts are right."""
    x = rng.integers(0, 2, B).astype(np.uint8)
    y = rng.integers(0, 2, B).astype(np.uint8)
    c0, c1 = keys.encrypt(x, 31000), keys.encrypt(y, 32000)
    outs = [engine.gate_batch(oracle.XOR, c0, c1) for _ in range(3)]
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    assert np.array_equal(keys.decrypt(outs[0]), x ^ y)
    idx = rng.choice(B, min(B, 6), replace=False)
    assert np.array_equal(outs[0][idx], oracle.gate_exact(keys, oracle.XOR, c0[idx], c1[idx]))


def test_async_batches_overlap_and_match(engine, oracle, keys, rng):
    """tfhe_b200_gate_batch_async: several batches in flight on the context's internal streams, then one sync; results equal
    the synchronous call.  301 gates per batch: above two gates per SM, so overlapping batches are cut into 4-gate CTAs (the last
    one ragged) while the synchronous call below deals the same gates evenly -- same bits either way."""
    import torch
    B, K = 301, 5
    ins, outs, want = [], [], []
    for k in range(K):
        x = rng.integers(0, 2, B).astype(np.uint8)
        y = rng.integers(0, 2, B).astype(np.uint8)
        a = torch.from_numpy(keys.encrypt(x, 40000 + 1000 * k).view(np.int32)).pin_memory()
        b = torch.from_numpy(keys.encrypt(y, 50000 + 1000 * k).view(np.int32)).pin_memory()
        o = torch.empty((B, n + 1), dtype=torch.int32).pin_memory()
        ins.append((a, b)); outs.append(o); want.append(1 - (x & y))
    engine.reserve(B)
    for k in range(K):
        engine.gate_batch_async(0, ins[k][0].numpy().view(np.uint32), ins[k][1].numpy().view(np.uint32), outs[k].numpy().view(np.uint32))
    engine.sync()
    for k in range(K):
        got = outs[k].numpy().view(np.uint32)
        assert np.array_equal(keys.decrypt(got), want[k]), k
    ref = engine.gate_batch(0, ins[2][0].numpy().view(np.uint32), ins[2][1].numpy().view(np.uint32))
    assert np.array_equal(ref, outs[2].numpy().view(np.uint32))


def test_batch_overlap_modes_same_bits(engine, keys, rng):
    """tfhe_b200_set_batch_overlap: the two ways a full batch is cut into CTAs (dealt evenly / 4-gate CTAs only) and the
    per-call default give the same ciphertext bits; bad modes are rejected."""
    from rustfhe_b200 import TfheError
    B = 601
    x = rng.integers(0, 2, B).astype(np.uint8)
    y = rng.integers(0, 2, B).astype(np.uint8)
    c0, c1 = keys.encrypt(x, 71000), keys.encrypt(y, 72000)
    try:
        outs = []
        for mode in (-1, 0, 1):
            engine.set_batch_overlap(mode)
            outs.append(engine.gate_batch(3, c0, c1))   # XOR
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
        assert np.array_equal(keys.decrypt(outs[0]), x ^ y)
        with pytest.raises(TfheError):
            engine.set_batch_overlap(2)
    finally:
        engine.set_batch_overlap(-1)


def test_exact_mode_three_key_slices(oracle, keys, rng):
    """tfhe_b200_set_key_slices: 1 = FFT64 (default), 2 = NTT with two 16-bit key slices, 3 = NTT with three 11-bit slices, exact
    in the WORST case (every slice sum < p/2 for any key and any digits).  On real keys all three modes give
    the same ciphertext bits; the adversarial key (all words 0x7FFFFFFF) against all digits -32 is exact by construction in the
    three-slice mode only.  Switching re-transforms the loaded key."""
    import rustfhe_b200 as R
    eng = R.DeviceEngine(0)
    try:
        default_mode = eng.stats()["key_slices"]
        eng.load_ksk(keys.ksk)
        eng.load_bk(keys.bk)
        B = 300
        x = rng.integers(0, 2, B).astype(np.uint8)
        y = rng.integers(0, 2, B).astype(np.uint8)
        c0, c1 = keys.encrypt(x, 61000), keys.encrypt(y, 62000)
        eng.set_key_slices(1)
        fft = eng.gate_batch(R.NAND, c0, c1)
        eng.set_key_slices(2)
        fast = eng.gate_batch(R.NAND, c0, c1)
        eng.set_key_slices(3)
        exact = eng.gate_batch(R.NAND, c0, c1)
        assert np.array_equal(keys.decrypt(exact), 1 - (x & y))
        assert np.array_equal(fast, exact)
        assert np.array_equal(fft, exact)
        idx = rng.choice(B, 8, replace=False)
        assert np.array_equal(exact[idx], oracle.gate_exact(keys, oracle.NAND, c0[idx], c1[idx]))
        few = eng.gate_batch(R.XOR, c0[:3], c1[:3])          # latency shape (cluster pair), three slices
        assert np.array_equal(few, oracle.gate_exact(keys, oracle.XOR, c0[:3], c1[:3]))
        trlwe = u32(rng, 3, 2, N)
        trlwe[1] = 0x7DF7C000                                 # all digits -32
        trgsw = u32(rng, 3, 6, 2, N)
        trgsw[1] = 0x7FFFFFFF                                 # every key slice at its extreme
        out = eng.external_product_batch(trgsw, trlwe)
        for g in range(3):
            ref = np.zeros(2 * N, np.uint32)
            oracle.lib().orc_external_product_exact(trgsw[g].reshape(-1), trlwe[g].reshape(-1), 0x02084000, ref)
            assert np.array_equal(out[g].reshape(-1), ref), g
        eng.set_key_slices(default_mode)
        assert np.array_equal(eng.gate_batch(R.NAND, c0[:200], c1[:200]), exact[:200])
        with pytest.raises(R.TfheError):
            eng.set_key_slices(4)
    finally:
        eng.close()


def test_fft64_mode_exact(oracle, keys, rng):
    """The FFT64 throughput kernel (blind_rotate_f64_kernel: f64 transform, exact rounding, one warp per gate, key ring in shared
    memory) against the exact-integer oracle: raw accumulators after 0, 1, 2 and 17 CMUX steps with the rounding edge cases of
    (b, a), whole-gate ciphertexts on a sample of a ragged batch (8 gates per CTA: 1185 = full CTAs + uneven dealing), the same
    bits run after run, and all decrypts right."""
    import rustfhe_b200 as R
    eng = R.DeviceEngine(0)
    try:
        eng.set_key_slices(1)
        eng.load_ksk(keys.ksk)
        eng.load_bk(keys.bk)
        sms = eng.stats()["sm_count"]
        B = 3 * sms + 13                                         # above 3 #SMs: the FFT64 throughput kernel, not waves of the latency kernel
        bits = rng.integers(0, 2, B).astype(np.uint8)
        lin = oracle.gate_linear(oracle.NAND, keys.encrypt(bits, 0), keys.encrypt(1 - bits, 1000))
        lin[0, 0] = 0                      # bbar = 0
        lin[3, 0] = 0xFFE00000             # bbar = 2047
        lin[4, 0] = 0x80000000             # bbar = 1024
        lin[4, 1] = 0xFFF00000             # abar rounds up to 2048 -> 0
        lin[4, 2] = 0x7FF00000             # abar = 1024
        for nsteps in (0, 1, 2, 17):
            out = eng.blind_rotate_batch(lin, nsteps)
            for g in (0, 1, 3, 4, 7, 8, B - 1):
                ref = np.zeros(2 * N, np.uint32)
                oracle.lib().orc_blind_rotate_exact(keys.exact_handle(), lin[g], 0x02084000, nsteps, ref)
                assert np.array_equal(out[g].reshape(-1), ref), (nsteps, g)
        B = 8 * sms + 1
        x = rng.integers(0, 2, B).astype(np.uint8)
        y = rng.integers(0, 2, B).astype(np.uint8)
        c0, c1 = keys.encrypt(x, 81000), keys.encrypt(y, 82000)
        outs = [eng.gate_batch(R.XOR, c0, c1) for _ in range(3)]
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
        assert np.array_equal(keys.decrypt(outs[0]), x ^ y)
        idx = np.concatenate([rng.choice(B, 6, replace=False), [0, B - 1]])
        assert np.array_equal(outs[0][idx], oracle.gate_exact(keys, oracle.XOR, c0[idx], c1[idx]))
        ops = rng.integers(0, 7, B).astype(np.uint8)            # a circuit level of mixed gates in one launch
        mixed = eng.gate_batch_mixed(ops, c0, c1)
        sel = np.flatnonzero(ops == R.XOR)
        assert np.array_equal(mixed[sel], outs[0][sel])
    finally:
        eng.close()


def test_fft64_tensor_memory_variant_same_bits(keys, rng, monkeypatch):
    """The opt-in K5FT variant (TFHE_B200_F64_TMEM=1: output spectra, digit planes and the rounded mask of a gate live in tensor
    memory, twelve gates per SM) returns the same ciphertext bits as the default kernel, on a batch with uneven dealing."""
    import rustfhe_b200 as R
    outs = []
    x = y = c0 = c1 = None
    for flag in ("0", "1"):
        monkeypatch.setenv("TFHE_B200_F64_TMEM", flag)
        eng = R.DeviceEngine(0)
        try:
            eng.set_key_slices(1)
            eng.load_ksk(keys.ksk)
            eng.load_bk(keys.bk)
            if x is None:
                B = 12 * eng.stats()["sm_count"] + 5
                x = rng.integers(0, 2, B).astype(np.uint8)
                y = rng.integers(0, 2, B).astype(np.uint8)
                c0, c1 = keys.encrypt(x, 91000), keys.encrypt(y, 92000)
            outs.append(eng.gate_batch(R.NAND, c0, c1))
            assert eng.stats()["gates_per_cta"] == (12 if flag == "1" else 8)
        finally:
            eng.close()
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(keys.decrypt(outs[1]), 1 - (x & y))


def test_fft64_two_warp_throughput_variant_same_bits(keys, rng, monkeypatch):
    """The opt-in K5F2 variant (TFHE_B200_F64_KERNEL=w2: one gate on two warps, three radix-8 passes per transform, six gates
    per SM) returns the same ciphertext bits as the default kernel, on a batch with uneven dealing."""
    import rustfhe_b200 as R
    outs = []
    x = y = c0 = c1 = None
    monkeypatch.setenv("TFHE_B200_F64_LATENCY", "0")
    for name in ("k5f", "w2"):
        monkeypatch.setenv("TFHE_B200_F64_KERNEL", name)
        eng = R.DeviceEngine(0)
        try:
            eng.set_key_slices(1)
            eng.load_ksk(keys.ksk)
            eng.load_bk(keys.bk)
            if x is None:
                B = 6 * eng.stats()["sm_count"] + 5
                x = rng.integers(0, 2, B).astype(np.uint8)
                y = rng.integers(0, 2, B).astype(np.uint8)
                c0, c1 = keys.encrypt(x, 93000), keys.encrypt(y, 94000)
            outs.append(eng.gate_batch(R.XOR, c0, c1))
            assert eng.stats()["gates_per_cta"] == (6 if name == "w2" else 8)
        finally:
            eng.close()
    assert np.array_equal(outs[0], outs[1])
    assert np.array_equal(keys.decrypt(outs[1]), x ^ y)


def test_gpu_against_committed_golden_vectors(engine, keys):
    """tests/golden/gate_vectors.npz was produced in the authoring container with the reference's own FFT library
    (tests/golden/make_golden.py); this compares the GPU output with the COMMITTED vectors directly, so the check does not
    depend on the oracle having been rebuilt identically on this box: whole NAND ciphertexts bit for bit against the
    exact-integer entries, decrypted bits and phases against the reference-FFT entries (P2: |e_gpu - e_ref| < 1/16), and
    the external product against both (bit-exact / within 2 ulp)."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gate_vectors.npz"))
    assert int(g["seed"]) == keys.seed
    c0, c1 = keys.encrypt(g["x"], 7000), keys.encrypt(g["y"], 7100)
    assert np.array_equal(c0[:, :8], g["c0_head"]) and np.array_equal(c1[:, :8], g["c1_head"])
    out = engine.gate_batch(0, c0, c1)
    assert np.array_equal(out, g["nand_exact"])
    assert np.array_equal(keys.decrypt(out), g["nand_bits"]) and np.array_equal(g["nand_bits"], 1 - (g["x"] & g["y"]))
    e_gpu = (keys.phase(out).astype(np.int64) - g["nand_ref_phase"].astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
    assert np.abs(e_gpu).max() < 2 ** 32 / 16
    xp = engine.external_product_batch(keys.bk[:12 * N].reshape(1, 6, 2, N), g["xp_trlwe"].reshape(1, 2, N))[0].reshape(-1)
    assert np.array_equal(xp, g["xp_exact"])
    d = (xp.astype(np.int64) - g["xp_ref"].astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
    assert np.abs(d).max() <= 2


def test_mixed_opcode_batch(engine, oracle, keys, rng):
    """tfhe_b200_gate_batch_mixed: one launch for gates of different kinds (a circuit level) == the per-opcode launches, bit for
    bit; in1 of NOT / COPY gates is ignored."""
    import rustfhe_b200 as R
    B = 210
    ops = rng.integers(0, 7, B).astype(np.uint8)
    x = rng.integers(0, 2, B).astype(np.uint8)
    y = rng.integers(0, 2, B).astype(np.uint8)
    c0, c1 = keys.encrypt(x, 71000), keys.encrypt(y, 72000)
    out = engine.gate_batch_mixed(ops, c0, c1)
    want_bits = {R.NAND: 1 - (x & y), R.AND: x & y, R.OR: x | y, R.XOR: x ^ y, R.NOT: 1 - x, R.COPY: x, R.ANDNY: (1 - x) & y}
    for op in range(7):
        idx = np.flatnonzero(ops == op)
        assert len(idx) > 0
        ref = engine.gate_batch(op, c0[idx], None if op in (R.NOT, R.COPY) else c1[idx])
        assert np.array_equal(out[idx], ref), op
        assert np.array_equal(keys.decrypt(out[idx]), want_bits[op][idx]), op
    with pytest.raises(R.TfheError):
        engine.gate_batch_mixed(np.full(B, 9, np.uint8), c0, c1)


def test_cpp_host_side_homnand_bench():
    """examples/homnand_bench (the reference's examples/homnand-bench.rs on include/tfhe_b200.hpp: truth tables of
    nand/and/or/xor/not/mux on fresh encryptions, then 1024 gates as one batch) decrypts everything right."""
    import subprocess
    from rustfhe_b200 import build as B
    B.build_example()
    r = subprocess.run([B.EXAMPLE], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    assert "all decryptions right" in r.stdout and "hom_nand_batch: 1024 gates" in r.stdout
