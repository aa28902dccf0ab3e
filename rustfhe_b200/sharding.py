"""Multi-GPU plumbing for the gate path: one process per GPU, keys replicated once, independent gate shards.

SURVEY.md section 8(e): every gate of a batch is independent and the only shared state is the read-only key material, so
the batch is cut into contiguous shards of ceil(B/g) gates, BK and KSK are broadcast ONCE (NCCL over NVLink on GPUs) and no
collective runs per gate.  The reference has no parallel path at all (single thread, F11); this module is the host logic
of ours and is backend-agnostic so that it is covered by world_size-2 `gloo` tests on CPU.
"""
import numpy as np


def shard_bounds(batch, world):
    """Contiguous shards of ceil(batch/world) gates: [(start, end)] per rank (trailing ranks may be empty)."""
    per = -(-batch // world) if batch > 0 else 0
    return [(min(r * per, batch), min((r + 1) * per, batch)) for r in range(world)]


def replicate_keys(bk_words, ksk_words, rank, world, device=None):
    """Broadcast the torus-domain bootstrapping key and the key-switching key from rank 0.

    Returns int32 torch tensors (bk, ksk) on `device` (CUDA for NCCL, None/cpu for gloo).  Rank 0 passes numpy uint32
    arrays, the other ranks pass None."""
    import torch
    import torch.distributed as dist
    from . import _capi as K
    bk = torch.empty(K.BK_WORDS, dtype=torch.int32, device=device)
    ksk = torch.empty(K.KSK_WORDS, dtype=torch.int32, device=device)
    if rank == 0:
        bk.copy_(torch.from_numpy(np.ascontiguousarray(bk_words, np.uint32).reshape(-1).view(np.int32)))
        ksk.copy_(torch.from_numpy(np.ascontiguousarray(ksk_words, np.uint32).reshape(-1).view(np.int32)))
    if world > 1:
        dist.broadcast(bk, 0)
        dist.broadcast(ksk, 0)
    return bk, ksk


def evaluate_sharded(gate_fn, in0, in1, rank, world, gather=True, ops=None):
    """Evaluate one batch of gates split across ranks.  `gate_fn(in0_shard, in1_shard) -> out_shard` is the per-rank
    engine call (tfhe_b200_gate_batch).  With gather=True every rank returns the whole output batch (one all_gather of
    2544 B per ciphertext -- the per-LEVEL exchange of a levelised circuit, never per gate)."""
    import torch
    import torch.distributed as dist
    B = len(in0)
    bounds = shard_bounds(B, world)
    s, e = bounds[rank]
    width = in0.shape[1]
    if e > s:
        args = (in0[s:e], None if in1 is None else in1[s:e])
        out_shard = gate_fn(*args) if ops is None else gate_fn(ops[s:e], *args)   # ops: per-gate opcodes of a mixed level
    else:
        out_shard = np.zeros((0, width), np.uint32)
    if not gather or world == 1:
        return out_shard
    per = bounds[0][1] - bounds[0][0]
    pad = np.zeros((per, width), np.uint32)
    pad[:e - s] = out_shard
    mine = torch.from_numpy(pad.view(np.int32))
    if dist.get_backend() == "nccl":   # NCCL moves device memory only; gloo (CPU tests) takes the host tensor as is
        mine = mine.cuda()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.concatenate([p.cpu().numpy().view(np.uint32)[:b[1] - b[0]] for p, b in zip(parts, bounds)], axis=0)
    return out


def evaluate_circuit_sharded(level_fn, netlist, inputs, rank, world, shard_min=75):
    """A levelised netlist with one process per GPU (the torch.distributed counterpart of tfhe_b200_group_circuit_run):
    every rank keeps the whole wire table; a level of at least `shard_min` gates is cut into contiguous shards (rustfhe_b200.
    circuit.level_plan), each rank evaluates its shard and ONE all_gather per level exchanges the level's outputs (2544 B per
    gate); narrower levels are evaluated by every rank on its own copy, without an exchange.
    `level_fn(ops, in0_rows, in1_rows) -> out_rows` is the per-rank engine call (tfhe_b200_gate_batch_mixed)."""
    import torch
    import torch.distributed as dist
    from . import circuit as Cq
    W = inputs.shape[1] if netlist.n_inputs else 636
    wires = np.zeros((netlist.n_wires, W), np.uint32)
    if netlist.n_inputs:
        wires[:netlist.n_inputs] = inputs
    for w, bit in netlist.consts.items():
        wires[w, 0] = 0x20000000 if bit else 0xE0000000
    sizes, ops, i0, i1, o = Cq._flatten_levels(netlist)
    plan = Cq.level_plan(sizes, world, shard_min)
    first = 0
    stats = {"sharded_levels": 0, "replicated_levels": 0}
    for width, (kind, what) in zip(sizes, plan):
        sl = slice(first, first + width)
        first += width
        if width == 0:
            continue
        if kind == "replicated":
            wires[o[sl]] = level_fn(ops[sl], wires[i0[sl]], wires[i1[sl]])
            stats["replicated_levels"] += 1
            continue
        f, c = what[rank]
        mine = level_fn(ops[sl][f:f + c], wires[i0[sl][f:f + c]], wires[i1[sl][f:f + c]]) if c else np.zeros((0, W), np.uint32)
        per = max(cnt for _, cnt in what)
        pad = np.zeros((per, W), np.uint32)
        pad[:c] = mine
        t = torch.from_numpy(pad.view(np.int32))
        if dist.get_backend() == "nccl":
            t = t.cuda()
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        rows = np.concatenate([p.cpu().numpy().view(np.uint32)[:cnt] for p, (_, cnt) in zip(parts, what)], axis=0)
        wires[o[sl]] = rows
        stats["sharded_levels"] += 1
    return (wires[netlist.outputs] if netlist.outputs else wires), stats

