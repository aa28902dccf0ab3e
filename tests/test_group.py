"""Device groups: one process drives several GPUs through the C ABI (include/tfhe_b200.h, rustfhe_b200/csrc/group.cu).
Keys are replicated inside the library by an NCCL broadcast; batches are sharded contiguously; results are bit-exact against
the exact-integer oracle and identical to a single-device context."""
import numpy as np
import pytest

N, n = 1024, 635


def test_group_fails_loudly_without_gpu():
    import torch
    import rustfhe_b200 as R
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(R.TfheError) as ei:
        R.DeviceGroup()
    assert ei.value.code == 2 and "no CPU fallback" in str(ei.value)


@pytest.mark.gpu
def test_group_one_device_matches_engine(engine, oracle, keys, rng):
    import rustfhe_b200 as R
    g = R.DeviceGroup([0])
    try:
        assert g.size == 1 and g.shard(10, 0) == (0, 10)
        g.load_ksk(keys.ksk)
        g.load_bk(keys.bk)
        B = 37
        x = rng.integers(0, 2, B).astype(np.uint8)
        y = rng.integers(0, 2, B).astype(np.uint8)
        c0, c1 = keys.encrypt(x, 81000), keys.encrypt(y, 82000)
        out = g.gate_batch(R.NAND, c0, c1)
        assert np.array_equal(out, engine.gate_batch(R.NAND, c0, c1))
        assert np.array_equal(keys.decrypt(out), 1 - (x & y))
        assert np.array_equal(out[:4], oracle.gate_exact(keys, oracle.NAND, c0[:4], c1[:4]))
        assert np.array_equal(keys.decrypt(g.gate_batch(R.NOT, c0)), 1 - x)
    finally:
        g.close()


@pytest.mark.gpu
def test_group_two_devices_bit_exact(engine, oracle, keys, rng):
    """>= 2 GPUs (run with gpurun --gpus 2): keys generated on device 0 and replicated by the library's ncclBroadcast; a
    ragged batch sharded over the devices equals the single-device result bit for bit and the exact oracle on a sample;
    every device has run a share."""
    import torch
    import rustfhe_b200 as R
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs two GPUs")
    g = R.DeviceGroup(list(range(min(ndev, 8))))
    try:
        g.keygen(keys.seed, keys.s0, keys.s1)          # same seeded keys as the oracle's (device keygen is bit-identical)
        B = 1531
        x = rng.integers(0, 2, B).astype(np.uint8)
        y = rng.integers(0, 2, B).astype(np.uint8)
        c0, c1 = keys.encrypt(x, 91000), keys.encrypt(y, 92000)
        g.reserve(B)
        out = g.gate_batch(R.XOR, c0, c1)
        assert np.array_equal(keys.decrypt(out), x ^ y)
        assert np.array_equal(out, engine.gate_batch(R.XOR, c0, c1))
        idx = rng.choice(B, 6, replace=False)
        assert np.array_equal(out[idx], oracle.gate_exact(keys, oracle.XOR, c0[idx], c1[idx]))
        for r in range(g.size):
            first, count = g.shard(B, r)
            assert g.ctx_stats(r)["last_batch"] == count
        # host-loaded keys take the same broadcast path
        g.load_ksk(keys.ksk)
        g.load_bk(keys.bk)
        assert np.array_equal(g.gate_batch(R.XOR, c0[:300], c1[:300]), out[:300])
    finally:
        g.close()


@pytest.mark.gpu
def test_group_circuit_one_device_matches_device_circuit(engine, keys, rng):
    """tfhe_b200_group_circuit_* on a group of one device (every level takes the replicated path): same output bits as the
    single-context device circuit, and the decrypted sum is right."""
    import rustfhe_b200 as R
    from rustfhe_b200 import circuit as Cq
    nl = Cq.ripple_carry_adder(8)
    bits = rng.integers(0, 2, nl.n_inputs).astype(np.uint8)
    cts = keys.encrypt(bits, 96000)
    dc = Cq.DeviceCircuit(engine, nl)
    want = dc.run(cts)
    dc.close()
    g = R.DeviceGroup([0])
    try:
        g.load_ksk(keys.ksk)
        g.load_bk(keys.bk)
        gc = Cq.GroupCircuit(g, nl, shard_min=1)
        got = gc.run(cts)
        assert gc.last["sharded_levels"] == 0 and gc.last["replicated_levels"] == gc.levels
        gc.close()
        assert np.array_equal(got, want)
        assert np.array_equal(keys.decrypt(got), nl.simulate(bits))
        # an output wire outside the table is refused with an error code (and the group stays usable)
        import dataclasses
        gc = Cq.GroupCircuit(g, nl)
        gc.netlist = dataclasses.replace(nl, outputs=[nl.n_wires + 5])
        with pytest.raises(R.TfheError) as ei:
            gc.run(cts)
        assert "out of range" in str(ei.value)
        gc.netlist = nl
        assert np.array_equal(gc.run(cts), want)
        gc.close()
        # constants only (a nander expression): trivial ciphertexts are set on the device, no inputs are uploaded
        ex = Cq.expr_to_netlist(Cq.parse_logic_expr("!(1&0)^(0|1)&1"))
        gc = Cq.GroupCircuit(g, ex)
        assert np.array_equal(keys.decrypt(gc.run()), ex.simulate([]))
        gc.close()
    finally:
        g.close()


@pytest.mark.gpu
def test_group_circuit_two_devices_sharded_levels_bit_exact(engine, keys, rng):
    """>= 2 GPUs: twelve 8-bit prefix adders side by side; with shard_min = 2 nearly every level is cut over the devices and its
    outputs exchanged by the library's NCCL broadcasts, with the default shard_min (one wave of the latency kernel) only the
    levels wider than that are.  Both runs equal the single-device circuit bit for bit."""
    import torch
    import rustfhe_b200 as R
    from rustfhe_b200 import circuit as Cq
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs two GPUs")
    nl = Cq.side_by_side(Cq.prefix_adder(8), 12)
    bits = rng.integers(0, 2, nl.n_inputs).astype(np.uint8)
    cts = keys.encrypt(bits, 97000)
    dc = Cq.DeviceCircuit(engine, nl)
    want = dc.run(cts)
    dc.close()
    assert np.array_equal(keys.decrypt(want), nl.simulate(bits))
    g = R.DeviceGroup(list(range(min(ndev, 8))))
    try:
        g.load_ksk(keys.ksk)
        g.load_bk(keys.bk)
        for smin in (2, 0):
            gc = Cq.GroupCircuit(g, nl, shard_min=smin)
            got = gc.run(cts)
            plan = Cq.level_plan(gc.sizes, g.size, gc.last["shard_min"])
            assert gc.last["sharded_levels"] == sum(1 for p in plan if p[0] == "sharded") > 0
            assert gc.last["sharded_levels"] + gc.last["replicated_levels"] == gc.levels
            got2 = gc.run(cts)                       # a second run reuses the device tables
            gc.close()
            assert np.array_equal(got, want) and np.array_equal(got2, want)
    finally:
        g.close()
