"""Pin the CPU oracle against every arithmetic KAT the reference's own tests hold for the path and
against the reference's real FFT library (oracle/_ref).  file:line citations are relative to /root/reference."""
import os
import subprocess
import sys

import numpy as np
import pytest

N = 1024


def U(xs):
    return np.array([int(x) & 0xFFFFFFFF for x in xs], dtype=np.uint32)


def rot(O, p, k):
    p = U(p)
    out = np.zeros_like(p)
    O.lib().orc_rotate(p, len(p), k, out)
    return out.astype(np.int32).tolist()


def test_rotate_kats(oracle):
    """utils/src/math.rs:75-84 (doctest) and 895-903 (polynomial_rotate)"""
    p = [1, 2, 3, 4, 5]
    assert rot(oracle, p, 1) == [-5, 1, 2, 3, 4]
    assert rot(oracle, p, -1) == [2, 3, 4, 5, -1]
    assert rot(oracle, p, 5) == [-1, -2, -3, -4, -5]
    assert rot(oracle, p, -4) == [5, -1, -2, -3, -4]
    assert rot(oracle, p, -8) == rot(oracle, p, 2)
    assert rot(oracle, p, 10) == p
    assert rot(oracle, p, 3) == [-3, -4, -5, 1, 2]
    assert rot(oracle, p, -3) == [4, 5, -1, -2, -3]


def cross(O, a, d):
    a, d = U(a), np.array(d, np.int32)
    out = np.zeros_like(a)
    O.lib().orc_negacyclic_mul_schoolbook(a, d, len(a), out)
    return out


def test_schoolbook_kats(oracle):
    """utils/src/math.rs:761-843 polynomial_cross, 845-864 polynomial_mul_add"""
    assert cross(oracle, [2, 3, 4], [4, 5, 6]).astype(np.int32).tolist() == [-30, -2, 43]
    h, q3 = 1 << 31, 3 << 30
    r = cross(oracle, [h, q3], [2, 3])                      # [0.5, 0.75] * [2, 3] = [0.75, 0.0]
    assert r.tolist() == [q3, 0]
    assert cross(oracle, [h], [1]).tolist() == [h]
    assert cross(oracle, [1 << 30, h], [1, 0]).tolist() == [1 << 30, h]
    # mul_add: [2,3,4]*[4,5,6] + [1,1,1] = [-29,-1,44]
    assert (cross(oracle, [2, 3, 4], [4, 5, 6]).astype(np.int32) + 1).tolist() == [-29, -1, 44]
    r = cross(oracle, [h, q3], [2, 3]) + U([1 << 29, 1 << 30])
    assert r.tolist() == [7 << 29, 1 << 30]               # [0.875, 0.25]


def dec_i32(O, x, l, bits):
    out = np.zeros(l, np.int32)
    O.lib().orc_decompose_scalar(x, l, bits, O.lib().orc_tested_decomp_mask(l, bits), out)
    return out.tolist()


def dec_u32(O, x, l, bits):
    out = np.zeros(l, np.uint32)
    O.lib().orc_decompose_u32_scalar(x, l, bits, out)
    return out.tolist()


def test_decomposition_kats(oracle):
    """utils/src/math.rs:1207-1273 decimal_decomposition and 866-893 polynomial_decomposition (these pin the OR-built mask
    of decomposition_i32, math.rs:582-591 -- NOT the production make_decomp_mask)"""
    O = oracle
    assert dec_u32(O, 0x80000000, 32, 1) == [1] + [0] * 31
    assert dec_i32(O, 0x80000000, 32, 1) == [-1] + [0] * 31
    assert dec_i32(O, 0x80000000, 8, 4) == [-8, 0, 0, 0, 0, 0, 0, 0]
    assert dec_i32(O, 0x80000000, 7, 4) == [-8, 0, 0, 0, 0, 0, 0]
    assert dec_u32(O, 0x80000001, 31, 1) == [1] + [0] * 29 + [1]
    assert dec_i32(O, 0x80000001, 31, 1) == [0] + [-1] * 30
    assert dec_i32(O, 0b000001_000010_000011_000000_000000_00, 3, 6) == [1, 2, 3]
    assert dec_i32(O, 0b000001_000010_000011_100000_000000_00, 3, 6) == [1, 2, 4]
    assert dec_i32(O, 0b011111_100000_100000_000000_100000_00, 3, 6) == [-32, -31, -32]
    assert dec_i32(O, 0x00000001, 2, 16) == [0, 1] and dec_i32(O, 0x00028000, 2, 16) == [3, -32768]


def test_production_mask_value_and_effect(oracle):
    """Torus32::make_decomp_mask(3,6) evaluates to 0x02084000 (rounding bit added twice, math.rs:546 + 548-551); the tested
    path builds 0x02082000.  The two give different digits on most inputs -- both are reproduced."""
    O = oracle
    assert O.lib().orc_make_decomp_mask(3, 6) == 0x02084000 == O.MASK_FAITHFUL
    assert O.lib().orc_tested_decomp_mask(3, 6) == 0x02082000 == O.MASK_TESTED
    rng = np.random.default_rng(3)
    xs = rng.integers(0, 2 ** 32, 4000, dtype=np.uint64)
    differ, maxerr_f, maxerr_t = 0, 0, 0
    a, b = np.zeros(3, np.int32), np.zeros(3, np.int32)
    for x in xs:
        O.lib().orc_decompose_scalar(int(x), 3, 6, O.MASK_FAITHFUL, a)
        O.lib().orc_decompose_scalar(int(x), 3, 6, O.MASK_TESTED, b)
        differ += a.tolist() != b.tolist()
        rec = lambda d: (int(d[0]) << 26) + (int(d[1]) << 20) + (int(d[2]) << 14)
        err = lambda d: abs(((rec(d) - int(x) + 2 ** 31) % 2 ** 32) - 2 ** 31)
        maxerr_f, maxerr_t = max(maxerr_f, err(a)), max(maxerr_t, err(b))
        assert all(-32 <= v < 32 for v in a) and all(-32 <= v < 32 for v in b)
    assert differ > 0.5 * len(xs)
    assert maxerr_t <= 2 ** 13 and 2 ** 13 < maxerr_f <= 2 ** 15


def test_torus_encoding_kats(oracle):
    """utils/src/math.rs:988-999 decimal_from_f32; tlwe.rs:181-194 binary2torus / torus2binary"""
    f = oracle.lib().orc_torus_from_f32
    assert f(0.5) == 1 << 31 and f(0.25) == 1 << 30 and f(0.125) == 1 << 29
    assert f(-0.5) == 1 << 31 and f(-0.25) == (1 << 30) + (1 << 31)
    assert f(1.0 / 8.0) == oracle.MU and f(-1.0 / 8.0) == 0xE0000000
    assert f(0.0) == 0 and f(1.0) == 0 and f(2.75) == 3 << 30


def test_tlwerep_linear_ops_kat(oracle):
    """hom_nand/src/tlwe.rs:302-326 tlwerep_op, re-expressed on the flat layout through the product's host mirror"""
    import rustfhe_b200 as R
    t = oracle.lib().orc_torus_from_f32
    l = np.zeros((1, 636), np.uint32)
    r = np.zeros((1, 636), np.uint32)
    l[0, :3] = [t(0.5), t(0.5), t(0.25)]
    r[0, :3] = [t(0.25), t(0.125), t(0.5)]
    assert R.TLWERep.add(l, r)[0, :3].tolist() == [t(0.75), t(0.625), t(0.75)]
    assert R.TLWERep.sub(l, r)[0, :3].tolist() == [t(0.25), t(0.375), t(0.75)]
    assert R.TLWERep.mul(l, 3)[0, :3].tolist() == [t(0.5), t(0.5), t(0.75)]
    assert R.TLWERep.mul(l, 0)[0, :3].tolist() == [0, 0, 0]
    assert R.TLWERep.mul(l, -1).tolist() == R.TLWERep.neg(l).tolist()
    assert R.TLWERep.logic_true()[0, 0] == 0x20000000 and R.TLWERep.logic_false()[0, 0] == 0xE0000000


def test_gate_linear_matches_reference_formulas(oracle, rng):
    """tfhe.rs:27-71 pre-combinations against the host mirror's TLWERep algebra"""
    import rustfhe_b200 as R
    x = rng.integers(0, 2 ** 32, (3, 636), dtype=np.uint64).astype(np.uint32)
    y = rng.integers(0, 2 ** 32, (3, 636), dtype=np.uint64).astype(np.uint32)
    T = R.TLWERep
    mu = T.trivial(0x20000000, 3)
    assert np.array_equal(oracle.gate_linear(oracle.NAND, x, y), T.sub(mu, T.add(x, y)))
    assert np.array_equal(oracle.gate_linear(oracle.AND, x, y), T.sub(T.add(x, y), mu))
    assert np.array_equal(oracle.gate_linear(oracle.OR, x, y), T.add(T.add(x, y), mu))
    assert np.array_equal(oracle.gate_linear(oracle.XOR, x, y), T.add(T.mul(T.add(x, y), 2), T.trivial(0x40000000, 3)))
    assert np.array_equal(oracle.gate_linear(oracle.NOT, x), T.neg(x))
    assert np.array_equal(oracle.gate_linear(oracle.ANDNY, x, y), T.sub(T.add(T.neg(x), y), mu))


def test_exact_ntt_equals_schoolbook(oracle, rng):
    """the exact layer's fast product (independent 2-prime CPU NTT + CRT) against the schoolbook definition"""
    for trial in range(4):
        a = rng.integers(0, 2 ** 32, N, dtype=np.uint64).astype(np.uint32)
        d = rng.integers(-32, 32, N).astype(np.int32)
        if trial == 3:
            a[:], d[:] = 0x80000000, -32
        o1, o2 = np.zeros(N, np.uint32), np.zeros(N, np.uint32)
        oracle.lib().orc_negacyclic_mul_schoolbook(a, d, N, o1)
        oracle.lib().orc_negacyclic_mul_ntt(a, d, o2)
        assert np.array_equal(o1, o2)


def test_sample_extract_all_indices(oracle, keys, rng):
    """hom_nand/src/trlwe.rs:178-205: extract(i) is an LWE sample of coefficient i of the TRLWE phase, for all 1024 i"""
    trlwe = rng.integers(0, 2 ** 32, 2 * N, dtype=np.uint64).astype(np.uint32)
    ph = np.zeros(N, np.uint32)
    oracle.lib().orc_trlwe_phase(keys.s1, trlwe, ph)
    out = np.zeros(N + 1, np.uint32)
    p1 = np.zeros(1, np.uint32)
    for i in range(N):
        oracle.lib().orc_sample_extract(trlwe, i, out)
        oracle.lib().orc_tlwe1_phase(keys.s1, out, 1, p1)
        assert p1[0] == ph[i], i


def test_key_switch_preserves_message(oracle, keys, rng):
    """hom_nand/src/tlwe.rs:346-396: identity_key_switch 1024 -> 635 keeps the decrypted bit"""
    for bit in (0, 1):
        a = rng.integers(0, 2 ** 32, N, dtype=np.uint64).astype(np.uint32)
        b = np.uint32((int(a[keys.s1 == 1].astype(np.uint64).sum()) + (0x20000000 if bit else 0xE0000000)) & 0xFFFFFFFF)
        lwe1 = np.concatenate([[b], a]).astype(np.uint32)
        out = np.zeros(636, np.uint32)
        oracle.lib().orc_key_switch(keys.ksk, lwe1, out)
        assert keys.decrypt(out)[0] == bit
        dig = np.zeros(N, np.uint16)
        oracle.lib().orc_ks_digits(lwe1, dig)
        assert np.array_equal(dig, ((a.astype(np.uint64) + 0x8000) >> 16).astype(np.uint16))


def test_reference_fft_roundtrip_and_product(oracle, rng):
    """the reference's own FFT (oracle/_ref): ifft->fft round trip is exact (utils/src/spqlios.rs:247-257 at N=1024) and
    Spqlios_poly_mul agrees with exact integers within 1 ulp"""
    if not oracle.ref_init():
        pytest.skip("oracle/_ref not built")
    a = np.zeros(N, np.uint32)
    a[1] = a[2] = 1
    out = np.zeros(N, np.uint32)
    oracle.lib().orc_ref_ifft_fft_roundtrip(a, out)
    assert np.array_equal(a, out)
    x = rng.integers(0, 2 ** 32, N, dtype=np.uint64).astype(np.uint32)
    d = rng.integers(0, 2, N).astype(np.int32)             # torus x binary as in math.rs:905-952
    ref, ex = np.zeros(N, np.uint32), np.zeros(N, np.uint32)
    oracle.lib().orc_ref_poly_mul(x, d.astype(np.uint32), ref)
    oracle.lib().orc_negacyclic_mul_schoolbook(x, d, N, ex)
    diff = (ref.astype(np.int64) - ex.astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
    assert np.abs(diff).max() <= 1


def test_spqlios_fft_test_kat_n16():
    """utils/src/spqlios.rs:243-276 fft_test step 1 at N=16 -- in a SEPARATE process because the reference latches 2/N of
    the first processor created (fft_processor_spqlios.cpp:110,158).  Step 2 is vacuous in the reference
    (tolerance 1000 on values in [0,1), F7); we record what the library really returns."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "libspqlios_ref.so")):
        pytest.skip("oracle/_ref not built")
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from oracle import oracle as O\n"
        "assert O.ref_init()\n"
        "a = np.zeros(16, np.uint32); a[1] = a[2] = 1; out = np.zeros(16, np.uint32)\n"
        "O.lib().orc_ref_roundtrip_n(16, a, out); assert np.array_equal(a, out), out\n"
        "m = np.zeros(16, np.uint32); O.lib().orc_ref_poly_mul_n(16, a, a, m)\n"
        "d = (m.astype(np.int64) - np.array([0,0,1,2,1]+[0]*11)); assert np.abs(d).max() <= 1, m\n"
        "print('ok')\n" % root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


def test_external_product_ref_vs_exact(oracle, keys, rng):
    """one external product: reference FFT vs exact integers differ by at most 1 ulp per coefficient and the
    TRGSW(1) (x) TRLWE(m) behavioural test of hom_nand/src/trgsw.rs:363-390 (phase preserved within 2e-3)"""
    if not oracle.ref_init():
        pytest.skip("oracle/_ref not built")
    per = 12 * N
    for i in (0, 5, 77):
        trgsw = keys.bk[per * i:per * (i + 1)]
        trlwe = rng.integers(0, 2 ** 32, 2 * N, dtype=np.uint64).astype(np.uint32)
        ex, rf = np.zeros(2 * N, np.uint32), np.zeros(2 * N, np.uint32)
        oracle.lib().orc_external_product_exact(trgsw, trlwe, oracle.MASK_FAITHFUL, ex)
        oracle.lib().orc_ref_external_product_torus(trgsw, trlwe, oracle.MASK_FAITHFUL, rf)
        diff = (rf.astype(np.int64) - ex.astype(np.int64) + 2 ** 31) % 2 ** 32 - 2 ** 31
        assert np.abs(diff).max() <= 1
        pin, pout = np.zeros(N, np.uint32), np.zeros(N, np.uint32)
        oracle.lib().orc_trlwe_phase(keys.s1, trlwe, pin)
        oracle.lib().orc_trlwe_phase(keys.s1, ex, pout)
        want = pin.astype(np.int64) * int(keys.s0[i])       # TRGSW(s0_i) (x) TRLWE(m) ~ TRLWE(s0_i * m)
        e = ((pout.astype(np.int64) - want + 2 ** 31) % 2 ** 32 - 2 ** 31) / 2.0 ** 32
        assert np.abs(e).max() < 2e-3


@pytest.mark.parametrize("op", ["NAND", "AND", "OR", "XOR", "NOT"])
def test_oracle_gate_truth_tables(oracle, keys, op):
    """hom_nand/src/tfhe.rs:147-279 / examples/homnand-bench.rs: decrypted truth tables, both oracle layers"""
    x = np.array([0, 1, 0, 1], np.uint8)
    y = np.array([0, 0, 1, 1], np.uint8)
    want = {"NAND": 1 - (x & y), "AND": x & y, "OR": x | y, "XOR": x ^ y, "NOT": 1 - x}[op]
    c0, c1 = keys.encrypt(x, 7000), keys.encrypt(y, 7100)
    code = getattr(oracle, op)
    ex = oracle.gate_exact(keys, code, c0, None if op == "NOT" else c1)
    assert np.array_equal(keys.decrypt(ex), want)
    if oracle.ref_init():
        rf = oracle.gate_ref(keys, code, c0, None if op == "NOT" else c1)
        assert np.array_equal(keys.decrypt(rf), want)
