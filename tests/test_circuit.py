"""Circuit front-end: nander grammar (nander/src/lib.rs:90-172), netlist levelisation, 32-bit ripple-carry adder (config 4).
CPU tests check the host logic against cleartext simulation; the GPU test evaluates the encrypted adder."""
import numpy as np
import pytest

from rustfhe_b200 import circuit as Cq


def test_parse_examples_from_the_repl_banner():
    """examples printed by nander_console (nander/src/main.rs:27-31): 1&1 => 1 ; !(1|0)$0 => 1 ; 1&1$0 => (1&1)$0"""
    assert Cq.eval_logic_expr_plain(Cq.parse_logic_expr("1&1")) == 1
    assert Cq.eval_logic_expr_plain(Cq.parse_logic_expr("!(1|0)$0")) == 1
    e = Cq.parse_logic_expr(" 1 & 1 $ 0 ")
    assert e.kind == "nand" and e.lhs.kind == "and" and Cq.eval_logic_expr_plain(e) == 1
    assert Cq.eval_logic_expr_plain(Cq.parse_logic_expr("!!1^1|0")) == 0      # ((!!1)^1)|0, no precedence
    assert Cq.eval_logic_expr_plain(Cq.parse_logic_expr("1)garbage")) == 1    # trailing input ignored like the reference


@pytest.mark.parametrize("bad,msg", [("(1&0", "braket is not closed"), ("1&", "invalid element. this is none"),
                                     ("a", "invalid element"), ("", "invalid element. this is none")])
def test_parse_errors(bad, msg):
    with pytest.raises(ValueError) as ei:
        Cq.parse_logic_expr(bad)
    assert str(ei.value) == msg


def test_adder_netlist_and_levels(rng):
    nl = Cq.ripple_carry_adder(32)
    assert len(nl.gates) == 5 + 31 * 9 and nl.n_inputs == 64 and len(nl.outputs) == 33
    assert all(op == Cq.NAND for op, *_ in nl.gates)
    for _ in range(20):
        x, y = int(rng.integers(0, 2 ** 32)), int(rng.integers(0, 2 ** 32))
        bits = [(x >> i) & 1 for i in range(32)] + [(y >> i) & 1 for i in range(32)]
        out = nl.simulate(bits)
        assert sum(int(b) << i for i, b in enumerate(out)) == x + y
    lv = nl.levels()
    widths = [sum(len(o) for (_, _, o) in l.values()) for l in lv]
    assert sum(widths) == len(nl.gates)
    assert widths[0] == 32                        # every x_i NAND y_i is independent of the carry chain
    assert len(lv) < 2 * 32 + 8                   # carry chain: ~2 levels per bit
    # a level only reads wires produced by earlier levels
    ready = set(range(nl.n_inputs)) | set(nl.consts)
    for l in lv:
        outs = set()
        for (i0, i1, o) in l.values():
            assert set(i0.tolist()) <= ready and set(i1.tolist()) <= ready
            outs |= set(o.tolist())
        ready |= outs


def test_expr_netlist_matches_plain(rng):
    for text in ("1&0|1^1", "!(1$1)&(0|!0)", "((1^1)^(1^0))$!(0&1)", "1"):
        e = Cq.parse_logic_expr(text)
        nl = Cq.expr_to_netlist(e)
        assert int(nl.simulate([])[0]) == Cq.eval_logic_expr_plain(e)


def test_evaluate_with_oracle_engine(oracle, keys):
    """the level-synchronous evaluator driven by the CPU oracle standing in for the device engine (host logic only)"""
    class OracleEngine:
        def gate_batch(self, op, a, b=None):
            return oracle.gate_exact(keys, op, a, b)
    nl = Cq.ripple_carry_adder(2)
    x, y = 3, 2
    bits = np.array([(x >> i) & 1 for i in range(2)] + [(y >> i) & 1 for i in range(2)], np.uint8)
    st = {}
    out = Cq.evaluate(OracleEngine(), nl, keys.encrypt(bits, 555), st)
    got = keys.decrypt(out)
    assert sum(int(b) << i for i, b in enumerate(got)) == x + y
    assert st["gates"] == 14 and sum(st["width_histogram"]) == 14


@pytest.mark.gpu
def test_encrypted_32bit_adder_on_gpu(engine, keys, rng):
    """config 4: two encrypted u32 (seed S+2), 284 NAND gates levelised; 33 output bits must equal x + y"""
    r = np.random.default_rng(0x5EED0001 + 2)
    x, y = int(r.integers(0, 2 ** 32)), int(r.integers(0, 2 ** 32))
    bits = np.array([(x >> i) & 1 for i in range(32)] + [(y >> i) & 1 for i in range(32)], np.uint8)
    nl = Cq.ripple_carry_adder(32)
    st = {}
    out = Cq.evaluate(engine, nl, keys.encrypt(bits, 31337), st)
    got = keys.decrypt(out)
    assert sum(int(b) << i for i, b in enumerate(got)) == x + y
    assert st["levels"] == len(st["width_histogram"])


@pytest.mark.gpu
def test_nander_expressions_on_gpu(engine, keys):
    class P:  # minimal `pros` carrying the engine, as TFHE does
        pass
    p = P()
    p.engine = engine
    for text in ("1&1", "!(1|0)$0", "1&1$0", "!(1$1)&(0|!0)^1"):
        e = Cq.parse_logic_expr(text)
        out = Cq.eval_logic_expr(p, e)
        assert int(keys.decrypt(out)[0]) == Cq.eval_logic_expr_plain(e), text


def test_prefix_adder_cleartext_and_shape():
    """Kogge-Stone adder on native gates: exact on random and corner operands, 11 levels and 451 gates for 32 bits."""
    from rustfhe_b200 import circuit as Cq
    nl = Cq.prefix_adder(32)
    rng = np.random.default_rng(5)
    cases = [(0, 0), (0xFFFFFFFF, 1), (0xFFFFFFFF, 0xFFFFFFFF), (0x80000000, 0x80000000), (0x55555555, 0xAAAAAAAB)]
    cases += [(int(rng.integers(0, 2 ** 32)), int(rng.integers(0, 2 ** 32))) for _ in range(50)]
    for x, y in cases:
        bits = [(x >> i) & 1 for i in range(32)] + [(y >> i) & 1 for i in range(32)]
        out = nl.simulate(bits)
        assert sum(int(b) << i for i, b in enumerate(out)) == x + y, (x, y)
    lv = nl.levels()
    assert len(lv) <= 12 and len(nl.gates) == 451
    assert len(nl.gates) < 700
    for k in (1, 2, 5, 8):
        n2 = Cq.prefix_adder(k)
        for x in range(min(2 ** k, 8)):
            for y in range(min(2 ** k, 8)):
                bits = [(x >> i) & 1 for i in range(k)] + [(y >> i) & 1 for i in range(k)]
                assert sum(int(b) << i for i, b in enumerate(n2.simulate(bits))) == x + y


@pytest.mark.gpu
def test_device_resident_circuit_matches_host_evaluator():
    """tfhe_b200_circuit_*: the levelised netlist lives on the device, a run is one launch pair per level with no host round
    trip; outputs are bit-identical to the host-table evaluator (and decrypt to x + y)."""
    import rustfhe_b200 as R
    from rustfhe_b200 import circuit as Cq
    sk = R.SecretKeys.generate(0x5EED0001)
    tfhe = R.TFHE.new_on_device(sk.s_key_tlwelv0, sk.s_key_tlwelv1, 0x5EED0001)
    try:
        for nl in (Cq.ripple_carry_adder(8), Cq.prefix_adder(16)):
            k = nl.n_inputs // 2
            x, y = 0xB7 & (2 ** k - 1), 0x5D3 & (2 ** k - 1)
            bits = np.array([(x >> i) & 1 for i in range(k)] + [(y >> i) & 1 for i in range(k)], np.uint8)
            cts = R.Cryptor.encrypto(R.TLWE, sk.s_key_tlwelv0, bits, seed=9, ct_index0=0)
            host = Cq.evaluate(tfhe.engine, nl, cts)
            dc = Cq.DeviceCircuit(tfhe.engine, nl)
            dev = dc.run(cts)
            dc.close()
            assert np.array_equal(dev, host)
            got = R.Cryptor.decrypto(R.TLWE, sk.s_key_tlwelv0, dev)
            assert sum(int(b) << i for i, b in enumerate(got)) == x + y
    finally:
        tfhe.close()


@pytest.mark.gpu
def test_device_circuit_rejects_racy_netlists(engine):
    """tfhe_b200_circuit_create: a level runs in place on one wire table with all its gates concurrent, so a wire written twice
    in a level, or read by one gate and written by another of the same level, is rejected (it would be a silent data race)."""
    import ctypes as C
    import rustfhe_b200 as R
    from rustfhe_b200 import _capi as K

    def create(level_sizes, ops, i0, i1, out, n_wires):
        h = C.c_void_p()
        sizes = (C.c_size_t * len(level_sizes))(*level_sizes)
        rc = engine._l.tfhe_b200_circuit_create(engine._ctx, len(level_sizes), sizes, K.ptr(np.asarray(ops, np.uint8)), K.ptr(np.asarray(i0, np.int32)),
                                                K.ptr(np.asarray(i1, np.int32)), K.ptr(np.asarray(out, np.int32)), n_wires, C.byref(h))
        if rc == 0:
            engine._l.tfhe_b200_circuit_destroy(engine._ctx, h)
        return rc
    assert create([2, 1], [R.NAND, R.AND, R.OR], [0, 0, 2], [1, 1, 3], [2, 3, 4], 5) == 0            # well formed
    assert create([2], [R.NAND, R.AND], [0, 0], [1, 1], [2, 2], 3) != 0                               # two gates write wire 2
    assert create([2], [R.NAND, R.AND], [0, 2], [1, 1], [2, 3], 4) != 0                               # wire 2 written and read in one level
    assert create([1], [R.NOT], [0], [0], [0], 1) != 0                                                # in place on its own input
    assert create([1], [R.NAND], [0], [5], [1], 3) != 0                                               # index out of range
    msg = engine._l.tfhe_b200_last_error(engine._ctx)
    msg = msg.decode() if isinstance(msg, bytes) else str(msg)
    assert "out of range" in msg


def test_side_by_side_and_level_plan(rng):
    """Host logic of the multi-GPU circuit evaluator: k disjoint copies of a netlist (levels k times as wide) simulate like k
    separate runs; the level plan shards a level of at least shard_min gates contiguously (every gate exactly once, the first
    width % world devices one gate more) and replicates the narrower ones."""
    base = Cq.prefix_adder(8)
    k = 5
    wide = Cq.side_by_side(base, k)
    assert wide.n_inputs == k * base.n_inputs and len(wide.gates) == k * len(base.gates)
    assert [sum(len(o) for (_, _, o) in lev.values()) for lev in wide.levels()] == \
           [k * sum(len(o) for (_, _, o) in lev.values()) for lev in base.levels()]
    bits = rng.integers(0, 2, (k, base.n_inputs)).astype(np.uint8)
    got = wide.simulate(bits.reshape(-1)).reshape(k, -1)
    for c in range(k):
        assert np.array_equal(got[c], base.simulate(bits[c]))
    sizes = [3, 75, 0, 1531, 74, 8]
    plan = Cq.level_plan(sizes, 8, 75)
    assert [p[0] for p in plan] == ["replicated", "sharded", "replicated", "sharded", "replicated", "replicated"]
    for w, (kind, what) in zip(sizes, plan):
        if kind == "replicated":
            assert what == w
        else:
            assert what[0][0] == 0 and all(what[r][0] + what[r][1] == what[r + 1][0] for r in range(7)) and what[7][0] + what[7][1] == w
            counts = [c for _, c in what]
            assert max(counts) - min(counts) <= 1 and counts == sorted(counts, reverse=True)
    assert all(p[0] == "replicated" for p in Cq.level_plan(sizes, 1, 1))


@pytest.mark.gpu
def test_circuit_level_pieces_match_whole_run(engine, keys, rng):
    """tfhe_b200_circuit_run_level_device / _scatter_level_device (what a group does per device): every level evaluated as two
    shards into a row buffer and then scattered gives the same wire table, bit for bit, as tfhe_b200_circuit_run_device."""
    import ctypes as C
    import torch
    import rustfhe_b200 as R
    from rustfhe_b200 import _capi as K
    nl = Cq.prefix_adder(16)
    bits = rng.integers(0, 2, nl.n_inputs).astype(np.uint8)
    cts = keys.encrypt(bits, 95000)
    dc = Cq.DeviceCircuit(engine, nl)
    try:
        W = K.n + 1
        wires = np.zeros((nl.n_wires, W), np.uint32)
        wires[:nl.n_inputs] = cts
        dev = torch.device("cuda", engine.device)
        st = torch.cuda.current_stream(dev)
        whole = torch.from_numpy(wires.view(np.int32)).to(dev)
        engine._ck(engine._l.tfhe_b200_circuit_run_device(engine._ctx, dc._h, C.c_void_p(whole.data_ptr()), C.c_void_p(st.cuda_stream)))
        nlev, nw, mx = C.c_size_t(), C.c_size_t(), C.c_size_t()
        assert engine._l.tfhe_b200_circuit_shape(dc._h, C.byref(nlev), C.byref(nw), C.byref(mx)) == 0
        assert (nlev.value, nw.value, mx.value) == (dc.levels, nl.n_wires, max(dc.sizes))
        pieces = torch.from_numpy(wires.view(np.int32)).to(dev)
        rows = torch.zeros((mx.value, W), dtype=torch.int32, device=dev)
        for l in range(nlev.value):
            g = C.c_size_t()
            assert engine._l.tfhe_b200_circuit_level_gates(dc._h, l, C.byref(g)) == 0 and g.value == dc.sizes[l]
            cut = g.value // 3
            for first, count in ((0, cut), (cut, g.value - cut)):
                engine._ck(engine._l.tfhe_b200_circuit_run_level_device(engine._ctx, dc._h, l, first, count, C.c_void_p(pieces.data_ptr()),
                                                                        C.c_void_p(rows.data_ptr() + first * W * 4), C.c_void_p(st.cuda_stream)))
            engine._ck(engine._l.tfhe_b200_circuit_scatter_level_device(engine._ctx, dc._h, l, C.c_void_p(rows.data_ptr()),
                                                                        C.c_void_p(pieces.data_ptr()), C.c_void_p(st.cuda_stream)))
        torch.cuda.synchronize()
        assert torch.equal(whole, pieces)
        got = keys.decrypt(whole[torch.as_tensor(nl.outputs, device=dev)].cpu().numpy().view(np.uint32))
        assert np.array_equal(got, nl.simulate(bits))
        assert engine._l.tfhe_b200_circuit_run_level_device(engine._ctx, dc._h, nlev.value, 0, 1, C.c_void_p(pieces.data_ptr()), None,
                                                            C.c_void_p(st.cuda_stream)) != 0          # level out of bounds
        assert engine._l.tfhe_b200_circuit_run_level_device(engine._ctx, dc._h, 0, 1, dc.sizes[0], C.c_void_p(pieces.data_ptr()), None,
                                                            C.c_void_p(st.cuda_stream)) != 0          # range past the level
    finally:
        dc.close()
