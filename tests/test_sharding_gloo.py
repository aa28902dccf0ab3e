"""world_size-2 `gloo` tests (CPU) of the multi-GPU host logic: one-off key replication and independent gate shards with no
per-gate collective.  The per-rank engine call is replaced by the CPU oracle here (tests may use
it); on the GPU box the same code path runs with NCCL and the CUDA engine (bench.py)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds():
    from rustfhe_b200.sharding import shard_bounds
    assert shard_bounds(1024, 8) == [(i * 128, (i + 1) * 128) for i in range(8)]
    assert shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert shard_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    assert shard_bounds(0, 2) == [(0, 0), (0, 0)]
    for B in (1, 7, 1024, 65537):
        for g in (1, 2, 4, 8):
            b = shard_bounds(B, g)
            assert b[0][0] == 0 and b[-1][1] == B and all(b[i][1] == b[i + 1][0] for i in range(g - 1))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from rustfhe_b200.sharding import evaluate_sharded, replicate_keys
        keys = O.Keys(0x5EED0001) if rank == 0 else None
        bk, ksk = replicate_keys(keys.bk if rank == 0 else None, keys.ksk if rank == 0 else None, rank, world)
        digest = (int(bk.to(torch.int64).sum()), int(ksk.to(torch.int64).sum()))
        # every rank rebuilds an oracle key object from the REPLICATED words (secret key only needed for the check)
        K = O.Keys.__new__(O.Keys)
        ref = O.Keys(0x5EED0001)
        K.seed, K.s0, K.s1 = ref.seed, ref.s0, ref.s1
        K.bk, K.ksk = bk.numpy().view(np.uint32).copy(), ksk.numpy().view(np.uint32).copy()
        K._exact = K._fourier = None
        B = 5  # ragged: shards of 3 and 2
        x = np.array([0, 1, 1, 0, 1], np.uint8)
        y = np.array([1, 1, 0, 0, 1], np.uint8)
        c0, c1 = ref.encrypt(x, 0), ref.encrypt(y, 100)
        out = evaluate_sharded(lambda a, b: O.gate_exact(K, O.NAND, a, b), c0, c1, rank, world, gather=True)
        ok = bool(np.array_equal(ref.decrypt(out), 1 - (x & y))) and out.shape == (B, 636)
        q.put((rank, digest, ok, out[:, :4].tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_key_replication_and_sharded_gates():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == res[1][1], "key replicas differ"
    assert res[0][2] and res[1][2], "sharded evaluation decrypted wrong"
    assert res[0][3] == res[1][3], "gathered outputs differ between ranks"


def _circuit_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rustfhe_b200 import circuit as Cq
        from rustfhe_b200.sharding import evaluate_circuit_sharded
        MU = 0x20000000
        calls = []

        def level_fn(ops, a, b):   # stand-in for tfhe_b200_gate_batch_mixed on TRIVIAL ciphertexts (word 0 = +-1/8, the rest 0)
            x, y = (a[:, 0] < 2 ** 31).astype(np.uint8), (b[:, 0] < 2 ** 31).astype(np.uint8)
            table = {Cq.NAND: 1 - (x & y), Cq.AND: x & y, Cq.OR: x | y, Cq.XOR: x ^ y, Cq.NOT: 1 - x}
            out = np.zeros_like(a)
            for k, op in enumerate(ops):
                out[k, 0] = MU if table[int(op)][k] else (2 ** 32 - MU)
            calls.append(len(ops))
            return out
        nl = Cq.side_by_side(Cq.prefix_adder(8), 6)
        r = np.random.default_rng(3)
        bits = r.integers(0, 2, nl.n_inputs).astype(np.uint8)
        cts = np.zeros((nl.n_inputs, 4), np.uint32)
        cts[:, 0] = np.where(bits == 1, MU, 2 ** 32 - MU)
        out, stats = evaluate_circuit_sharded(level_fn, nl, cts, rank, world, shard_min=40)
        got = (out[:, 0] < 2 ** 31).astype(np.uint8)
        sizes = [sum(len(o) for (_, _, o) in lev.values()) for lev in nl.levels()]
        plan = Cq.level_plan(sizes, world, 40)
        want_calls = [w if kind == "replicated" else what[rank][1] for w, (kind, what) in zip(sizes, plan)]
        q.put((rank, bool(np.array_equal(got, nl.simulate(bits))), stats, calls == want_calls, out[:, 0].tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_circuit_levels():
    """evaluate_circuit_sharded (the one-process-per-GPU form of the group circuit): wide levels are cut over the ranks and their
    outputs exchanged by one all_gather per level, narrow levels are evaluated by every rank; both ranks end with the same,
    correct outputs, and each rank's engine saw exactly its shards."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + (os.getpid() % 40)
    procs = [ctx.Process(target=_circuit_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] and res[1][1], "sharded circuit evaluated wrong"
    assert res[0][2] == res[1][2] and res[0][2]["sharded_levels"] > 0 and res[0][2]["replicated_levels"] > 0
    assert res[0][3] and res[1][3], "a rank evaluated gates outside its shards"
    assert res[0][4] == res[1][4], "outputs differ between ranks"
