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
